#include "/root/reference/BitNetMCU_model_cnn.h"
