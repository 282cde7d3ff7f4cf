/* codec2_fdmdv.h -- the one constant the FSK tools take from this header: the scaling of 16-bit samples
 * (utils/codec2_fdmdv.h:67; fsk_demod.c:287,305-306 divide by it). */
#ifndef B200_COMPAT_CODEC2_FDMDV_H
#define B200_COMPAT_CODEC2_FDMDV_H
#include "comp.h"
#define FDMDV_SCALE 1000
#endif
