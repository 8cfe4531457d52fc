// present only so that callers including the reference's header of this name keep compiling
#ifndef SUPER4PCS_B200_UTILS_DISABLEWARNINGS_H_
#define SUPER4PCS_B200_UTILS_DISABLEWARNINGS_H_
#endif
