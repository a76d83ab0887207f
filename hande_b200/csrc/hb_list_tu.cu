// hande_b200: the list kernels (hb_list.cuh) of ONE bit-string width, selected with -DHB_TU_W=<1..4|32>; the W = 32 unit
// (the wide layout) is compiled with -DHB_OCC16.
#include "hb_list.cuh"

#if !defined(HB_TU_W)
#error "compile with -DHB_TU_W=<1..4|32>"
#endif
#define HB_CAT2_(a, b) a##b
#define HB_CAT2(a, b) HB_CAT2_(a, b)

const ListOps* HB_CAT2(hb_list_ops_w, HB_TU_W)() { return list_ops<HB_TU_W>(); }
