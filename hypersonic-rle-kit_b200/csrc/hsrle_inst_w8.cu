#define HSRLE_INST_W 8
#include "hsrle_inst.cuh"
