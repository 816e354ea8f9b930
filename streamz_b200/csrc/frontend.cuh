// Internal interface of the front-end kernels (frontend.cu) used by the C-ABI layer (capi.cu).
#pragma once
#include <cstdint>
#include <vector>

#include "common.cuh"

namespace szb {

// One unit of the extraction work queue: windows [w_begin, w_end) of a clip that has n_total windows.
struct Segment {
    unsigned long long pcm_off;  // sample offset of the clip's first 44.1 kHz sample in the PCM buffer
    unsigned long long out_row;  // feature row of the clip's window 0
    uint32_t n_total;
    uint32_t w_begin;
    uint32_t w_end;
    uint32_t pad;                // fused resampling: input samples of the clip at its original rate
};

szb_status upload_frontend_tables();
// appends the segments of clips [clip_begin, clip_end) to `segs`
void build_segments(const uint64_t* clip_off44, const uint64_t* win_off, uint32_t clip_begin, uint32_t clip_end, int sm_count,
                    std::vector<Segment>& segs, const uint64_t* clip_n_in = nullptr);
bool fused_resample_supported(uint32_t rate);
szb_status upload_segments(szb_ctx* ctx, const std::vector<Segment>& segs, uint32_t n_queues);
szb_status launch_extract(szb_ctx* ctx, const int16_t* d_pcm44, size_t seg_begin, size_t n_segs, uint32_t queue, float* d_feats,
                          bool aligned16, uint32_t fused_rate = 0);
szb_status launch_resample(szb_ctx* ctx, const int16_t* d_in, const uint64_t* d_in_off, const uint64_t* d_out_off,
                           uint32_t n_clips, uint64_t max_out, uint32_t rate, int16_t* d_out);
szb_status launch_augment(szb_ctx* ctx, const int16_t* d_in, uint64_t n, uint64_t shift, float gain, float noise_level, uint64_t key,
                          int16_t* d_out);
szb_status launch_downmix(szb_ctx* ctx, const int16_t* d_in, uint64_t n_in, uint32_t ch, int16_t* d_out, uint64_t n_out);

}  // namespace szb
