// oracle/shim/shim_retype.h — last line of every shim header: from here on (i.e. in the reference's
// own code only) `float` means `double` when the fp64 variant is being built.
#ifdef EKF_SHIM_DOUBLE
#ifndef float
#define float double
#endif
#endif
