// cr_filter.cuh -- executable / bitmap pre-filters (src/cr-filter.c, filter_x86_elf.c, filter_x86_pe.c, filter_bmp.c).
#pragma once
#include "cr_common.cuh"
struct LzChain;
struct FilterHost {
    void reset() {}
    void release() {}
    int run_window(LzChain&, const uint8_t*, uint8_t*, uint64_t, const std::vector<uint64_t>&, const std::vector<uint32_t>&, std::vector<uint8_t>&, int&) { return CRGPU_ERR_UNSUPPORTED; }
};
