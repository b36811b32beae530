#define HSRLE_INST_W 2
#include "hsrle_inst.cuh"
