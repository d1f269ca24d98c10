#include "algo/vnl_shim_algo.h"
