#define HSRLE_INST_W 3
#include "hsrle_inst.cuh"
