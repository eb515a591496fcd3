// Internal interface between the C ABI (st_abi.cu) and the kernels (st_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace st {

enum Mode : int { MODE_STEP = 0, MODE_RESET = 1, MODE_OBSERVE = 2 };

constexpr int kWarpsPerCta = 8;       // image modes: one env per warp, 8 envs per CTA
#ifndef ST_RAM_WARPS
#define ST_RAM_WARPS 4
#endif
constexpr int kRamWarpsPerCta = ST_RAM_WARPS;  // ram mode
constexpr int kImage = 84;            // ref:426: _observation always renders at 84
constexpr int kStateWords = 15;

// Everything a launch needs, passed by value (__grid_constant__).
struct Params {
    // engine configuration (StConfig, pre-digested on the host)
    int W, H;
    int lock_mod;  // max(lock_delay,0)+1, ref:175
    int step_reset, auto_reset;
    int reward_step, pen_height, pen_height_inc, adv_clears, high_scoring, pen_holes, pen_holes_inc;
    uint32_t fullmask;
    int col_words;  // 32-bit words per board column in the env record: 1 (H <= 31) or 2
    int stride;     // bytes per env record
    uint32_t seed_lo, seed_hi;
    long long env_id_base;
    // observation geometry (closed form of ref:76-114 at size 84)
    int obs_elems;
    int pitch, gap, inner_v, inner_h, pad_top, pad_left;
    uint32_t inv_h20;   // ceil(2^20 / H)      : i / H      for i < 2048
    uint32_t inv_hq20;  // ceil(2^20 / (H/4))  : q / (H/4)  when H % 4 == 0
    uint32_t inv_sw20;  // ceil(2^20 / (stride/4)) : word index -> record, thread-per-env kernel
    uint32_t inv_nq32;  // ceil(2^32 / (W*H/4)) : float4 slot of a group's block -> env, thread-per-env kernel
    uint32_t inv_w20;   // ceil(2^20 / W)      : (env, column) item -> env, thread-per-env kernel
    // buffers
    unsigned char *state;
    const uint8_t *actions;
    void *obs;       // float32 [n][obs_elems], or uint8 when obs_u8
    int obs_u8;
    void *term_obs;  // optional [n][obs_elems]: terminal observation of auto-reset envs
    float *reward;
    uint8_t *done;
    int32_t *info;
    const uint8_t *mask;
    const uint8_t *queue;
    int queue_len;
    int *err;
    unsigned long long *stats;
    long long n;
    int T;
    long long obs_t_stride, info_t_stride;
    int mode;
    int draw_piece;
    int tpe_epw;     // thread-per-env kernel: envs per group (one group = one warp pass)
    int tpe_l2;      // thread-per-env kernel: L2 eviction hints (bit 0: outputs evict_first, bit 1: records evict_last)
    int tpe_nrec;    // thread-per-env kernel: record buffers per warp (2 = next group prefetched; capped grids)
    int tpe_staged;  // thread-per-env kernel: observations through a shared-memory block + TMA bulk store
    int tpe_sync;    // thread-per-env kernel: CTA barrier between the engine and the store phases (one CTA per SM)
};

cudaError_t launch_main(const Params &p, int obs_type, cudaStream_t stream);
cudaError_t launch_init(const Params &p, cudaStream_t stream);
cudaError_t launch_get_state(const Params &p, uint8_t *boards, int32_t *scalars, cudaStream_t stream);
cudaError_t launch_set_state(const Params &p, const uint8_t *boards, const int32_t *scalars, cudaStream_t stream);
cudaError_t launch_render(const Params &p, int size, uint8_t *out, cudaStream_t stream);
const char *step_kernel_name(const Params &p, int obs_type);
unsigned long long launch_count();

}  // namespace st
