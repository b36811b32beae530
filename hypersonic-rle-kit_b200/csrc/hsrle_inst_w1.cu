#define HSRLE_INST_W 1
#include "hsrle_inst.cuh"
