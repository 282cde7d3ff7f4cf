/* dft_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle.h). placeholder, filled in with the dft_detect restatement. */
#include "oracle.h"
