#include "rrel_muset_obj.h"
