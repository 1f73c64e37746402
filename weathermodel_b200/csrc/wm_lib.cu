// wm_lib.cu -- single translation unit of libwm_b200.so (keeps g_wm_dev_error a single definition and
// lets nvcc inline across files). Build: see weathermodel_b200/build.py.
#include "wm_gemm.cu"
#include "wm_elementwise.cu"
#include "wm_attn.cu"
#include "wm_yield.cu"
#include "wm_encoder.cu"
#include "wm_api.cu"
