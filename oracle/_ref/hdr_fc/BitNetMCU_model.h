#include "/root/reference/BitNetMCU_model_fc.h"
