// conv.cu <-> conv_tc.cu: the tensor-core path of ibinn_conv_forward
#pragma once
#include "common.cuh"

bool conv_tc_supported(const ibinn_conv_args& a);
int64_t conv_tc_partials_floats(const ibinn_conv_args& a);
int64_t conv_tc_wprep_floats(const ibinn_conv_args& a);
int conv_tc_launch(const ibinn_conv_args& a, cudaStream_t st);

bool wgrad_tc_supported(const ibinn_wgrad_args& a);
int64_t wgrad_tc_workspace_floats(const ibinn_wgrad_args& a);
int wgrad_tc_launch(const ibinn_wgrad_args& a, float* workspace, cudaStream_t st);
