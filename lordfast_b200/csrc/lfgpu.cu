/* liblfgpu.so: the one translation unit of the product library (CUDA only, sm_100a). */
#include "lf_pipeline.inl"
