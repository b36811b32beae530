#define HSRLE_INST_W 4
#include "hsrle_inst.cuh"
