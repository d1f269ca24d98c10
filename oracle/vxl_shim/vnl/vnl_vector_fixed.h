#include "vnl_shim_core.h"
