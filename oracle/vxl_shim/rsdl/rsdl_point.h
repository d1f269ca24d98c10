#include "rsdl_kd_tree.h"
