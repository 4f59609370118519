// cuemu shim for <math_constants.h> (tests only).
#pragma once
#include <math.h>
#define CUDART_NAN (__builtin_nan(""))
#define CUDART_NAN_F (__builtin_nanf(""))
#define CUDART_INF (__builtin_inf())
#define CUDART_INF_F (__builtin_inff())
#define CUDART_PI_F 3.141592654f
