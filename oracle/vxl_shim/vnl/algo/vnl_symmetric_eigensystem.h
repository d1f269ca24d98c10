#include "vnl_shim_algo.h"
