#include "vcl_shim_all.h"
