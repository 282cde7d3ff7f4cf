/* comp.h -- the complex sample type of the codec2 FSK modem API (source-compatible stand-in for utils/comp.h:33-36 of
 * radiosonde_auto_rx, so that programs written against the reference headers compile against include/ unchanged). */
#ifndef B200_COMPAT_COMP_H
#define B200_COMPAT_COMP_H
#ifndef __COMP__
#define __COMP__
typedef struct { float real, imag; } COMP;
#endif
#endif
