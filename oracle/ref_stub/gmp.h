// Empty stand-in for <gmp.h>: the reference includes it (algebra_msm_VariableBaseMSM.cu:15) but calls no GMP function
// (cgbn's mpz backend is commented out, include/cgbn/cgbn.h:67-83).  Only the runtime libgmp.so.10 exists in this image.
#ifndef OZK_REF_STUB_GMP_H
#define OZK_REF_STUB_GMP_H
#endif
