// nafgpu.cu — single translation unit of libnafgpu.so (kernels in one TU: no relocatable device code needed).
#include "api.cu"
#include "naf_dec.cu"
#include "naf_enc.cu"
