#define HSRLE_INST_W 6
#include "hsrle_inst.cuh"
