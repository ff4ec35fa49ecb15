/* oracle GSL shim (test infrastructure): forwards to gsl_shim.h */
#include "../gsl_shim.h"
