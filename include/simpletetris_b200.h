/*
 * simpletetris_b200.h — C ABI of the B200-native batched SimpleTetris step path.
 *
 * This is the drop-in boundary: everything the reference does between
 * `TetrisEnv.step/reset` and its NumPy results (reference:
 * gym_simpletetris/envs/tetris_env.py, cited as ref:LINE) happens behind these
 * entry points, in hand-written sm_100a CUDA kernels.  Plain pointers and
 * sizes only; no torch types.  All `st_*` launch functions are asynchronous on
 * the CUDA stream passed in (`stream` is a `cudaStream_t`, 0 = legacy default
 * stream) and never allocate or free caller memory.  Return value: 0 on
 * success, otherwise a non-zero code (a `cudaError_t` or one of ST_E_*);
 * `st_last_error()` gives the message for the calling thread.
 *
 * Ownership: the caller (the Python host keeps them in torch tensors) owns
 * state, observation, reward, done, info, action, queue, stats and error
 * buffers.  The `st_host_*` family is the exception: a self-contained handle
 * that owns its device buffers and takes/returns HOST memory, for callers with
 * no device allocator of their own (the single-env `TetrisEnv` facade, and the
 * end-to-end benchmark).
 */
#ifndef SIMPLETETRIS_B200_H
#define SIMPLETETRIS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ST_API __attribute__((visibility("default")))
#else
#define ST_API
#endif

#define ST_ABI_VERSION 1
#define ST_MAX_WIDTH 32   /* a board row is at most one 32-bit word in registers */
#define ST_MAX_HEIGHT 63  /* a board column is one 32-bit word in the env record (height <= 31) or two */
#define ST_STATE_WORDS 15 /* scalar words at the head of every env record */
#define ST_INFO_WORDS 15  /* int32 per env in the info output */
#define ST_UNPACKED_WORDS 18
#define ST_STATS_WORDS 4

#define ST_OBS_RAM 0       /* ref:421-424: float32 [W][H] of 0/1 */
#define ST_OBS_GRAYSCALE 1 /* ref:426-431 + ref:76-114: float32 [84][84] of 0/128/190 */
#define ST_OBS_RGB 2       /* ref:433 + ref:117-122: float32 [84][84][3] */

#define ST_E_INVALID 1001  /* bad argument / unsupported geometry */
#define ST_E_NODEVICE 1002

/* sticky bits OR-ed by the kernels into *error_flag (device int32) */
#define ST_ERR_QUEUE_EXHAUSTED 1 /* injected piece queue ran out (wrapped around) */
#define ST_ERR_BAD_ACTION 2      /* action outside 0..6 (ref:245 raises KeyError) -> treated as idle */
#define ST_ERR_NO_PIECE 4        /* step before the first reset (ref:170-172: shape is None) -> no-op */

/*
 * Replaces the constructor arguments of ref:343-357 (TetrisEnv.__init__) and
 * ref:126-137 (TetrisEngine.__init__).  `render_mode` has no effect on the step
 * path (ref:362 stores it, nothing reads it).  Flags are 0/1.
 */
typedef struct StConfig {
    int32_t width;        /* ref:344, 1..ST_MAX_WIDTH */
    int32_t height;       /* ref:345, 1..ST_MAX_HEIGHT */
    int32_t obs_type;     /* ref:346, ST_OBS_* */
    int32_t extend_dims;  /* ref:347: trailing 1 in the shape only; bytes identical */
    int32_t lock_delay;   /* ref:356 / ref:175: counter modulo max(lock_delay,0)+1 */
    int32_t step_reset;   /* ref:357 / ref:248-249 */
    int32_t reward_step;              /* ref:256 */
    int32_t penalise_height;          /* ref:286-287 */
    int32_t penalise_height_increase; /* ref:288-292 */
    int32_t advanced_clears;          /* ref:266-269 */
    int32_t high_scoring;             /* ref:270-272 */
    int32_t penalise_holes;           /* ref:294-295 */
    int32_t penalise_holes_increase;  /* ref:296-297 */
    int32_t auto_reset;   /* 1: VecEnv semantics (gym<=0.25): on done, clear() (ref:306-315) in the same
                             call and return the reset observation; 0: reference single-env semantics */
    int32_t device;       /* CUDA ordinal that owns every device pointer passed with this config */
    int32_t obs_u8;       /* 0: observations are float32, as the reference returns them (ref:400) — the parity mode.
                             1: the same values (0/1, 0/128/190) stored as uint8, 4x less observation traffic; an
                             extension for image pipelines, benchmarked separately, never reported as the parity mode */
    uint64_t seed;        /* Philox key of the piece stream (replaces the global `random`, ref:2,187) */
    int64_t env_id_base;  /* global id of env 0 of this shard; stream of env e is keyed by base+e */
} StConfig;

/* Optional device-side extras; any pointer may be NULL. */
typedef struct StAux {
    const uint8_t *piece_queue; /* [n][queue_len] piece ids 0..6 (ref:19 order); replaces _choose_shape
                                   (ref:183-191) for parity runs: the k-th piece of an env's lifetime is
                                   queue[e][k] */
    int32_t queue_len;
    int32_t reserved;
    int32_t *error_flag;        /* device int32, sticky OR of ST_ERR_* */
    unsigned long long *stats;  /* device u64[ST_STATS_WORDS]: episodes, sum(time), sum(lines_cleared),
                                   sum(score) accumulated at every done */
    void *terminal_obs;         /* [n][st_obs_elems], same dtype as obs: with auto_reset, the observation the
                                   reference's step() returned at the terminal step (ref:301-304) is written
                                   here for the envs that are done in this call (gym <= 0.25 vector envs put it
                                   in info["terminal_observation"]); rows of other envs are left untouched */
} StAux;

/* ---- sizes ------------------------------------------------------------------ */
/* Bytes of one env record in `state`: 15 int32 scalars (packed piece, lock-delay counter, time, score,
 * lines_cleared, holes, piece_height, deaths, shape_counts[7] — the engine attributes of ref:165-181)
 * followed by `width` board columns (bit y of column x = cell (x, y), the [x][y] order of ref:140), one 32-bit word
 * each when height <= 31, else two (low word first).  100 B at 10x20, 220 B at 20 wide x 40 high. */
ST_API int64_t st_state_stride(const StConfig *cfg);
/* float32 elements of one env's observation (ref:381-392): W*H, 7056 or 21168. */
ST_API int64_t st_obs_elems(const StConfig *cfg);

/* ---- the step path ------------------------------------------------------------ */
/* TetrisEngine.__init__ state (ref:140,165-181): empty board, time = score = -1, no piece. */
ST_API int st_init(const StConfig *cfg, void *state, int64_t n, void *stream);

/* TetrisEnv.reset (ref:405-411) -> TetrisEngine.clear (ref:306-315) for every env (mask == NULL) or the
 * envs with mask[e] != 0.  Writes the reset observation (empty board, piece not drawn) of those envs
 * into obs (may be NULL). */
ST_API int st_reset(const StConfig *cfg, void *state, const uint8_t *mask, void *obs, const StAux *aux,
             int64_t n, void *stream);

/* TetrisEnv.step (ref:397-403) -> TetrisEngine.step (ref:243-304) + _observation (ref:413-433) for n envs.
 *   actions [n] uint8 (ids of ref:152-160);  obs [n][st_obs_elems] float32 (uint8 if cfg->obs_u8);  reward [n] float32;
 *   done [n] uint8;  info [n][ST_INFO_WORDS] int32 or NULL: piece id, lock-delay counter, time, score,
 *   lines_cleared, holes, piece_height, deaths, shape_counts[7] (get_info, ref:232-241), taken after the
 *   step and BEFORE any auto-reset. */
ST_API int st_step(const StConfig *cfg, void *state, const uint8_t *actions, void *obs, float *reward,
            uint8_t *done, int32_t *info, const StAux *aux, int64_t n, void *stream);

/* T consecutive TetrisEnv.step calls (ref:397-403) in one launch: the caller loop of README.md:43-51 moved on the
 * device (state stays in registers between steps).  actions [T][n];
 * reward/done [T][n]; obs and info advance by obs_t_stride / info_t_stride ELEMENTS per step (0 = every
 * step overwrites the same [n][...] buffer, n*elems = a rollout buffer). */
ST_API int st_step_many(const StConfig *cfg, void *state, const uint8_t *actions, int32_t T, void *obs,
                 int64_t obs_t_stride, float *reward, uint8_t *done, int32_t *info, int64_t info_t_stride,
                 const StAux *aux, int64_t n, void *stream);

/* _observation of the current state without stepping (TetrisEngine.render, ref:317-321, when
 * draw_piece != 0; the bare board otherwise). */
ST_API int st_observe(const StConfig *cfg, const void *state, int32_t draw_piece, void *obs, int64_t n, void *stream);

/* TetrisEnv.render('rgb_array') (ref:458-462): engine.render() -> convert_grayscale(obs, size) ->
 * convert_grayscale_rgb, uint8 out [n][size][size][3]; the reference uses size 160. */
ST_API int st_render(const StConfig *cfg, const void *state, int32_t draw_piece, int32_t size, uint8_t *out,
                     int64_t n, void *stream);

/* ---- state injection / inspection (tests, checkpoints) -------------------------- */
/* What a test of the reference does by reading / assigning engine.board (ref:140), engine.shape / anchor /
 * shape_name (ref:170-172,196-200), the counters of ref:165-169,173, _lock_delay (ref:176) and shape_counts (ref:181).
 * boards [n][W][H] uint8 (ref board[x,y]);  scalars [n][ST_UNPACKED_WORDS] int32: piece id (7 = none),
 * rotation (number of rotate_left applications, ref:22-26), anchor x, anchor y, lock-delay counter, time,
 * score, lines_cleared, holes, piece_height, deaths, shape_counts[7].  Either pointer may be NULL. */
ST_API int st_get_state(const StConfig *cfg, const void *state, uint8_t *boards, int32_t *scalars, int64_t n, void *stream);
ST_API int st_set_state(const StConfig *cfg, void *state, const uint8_t *boards, const int32_t *scalars, int64_t n, void *stream);

/* ---- host-buffer handle (owns its device memory; synchronous) -------------------- */
typedef struct StHostEnv StHostEnv;
/* TetrisEnv.__init__ (ref:343-392) for n envs: allocates state/obs/... on cfg->device and runs st_init.  NULL on
 * failure (st_last_error() says why). */
ST_API StHostEnv *st_host_create(const StConfig *cfg, int64_t n);
ST_API void st_host_destroy(StHostEnv *h);
/* Replaces engine._choose_shape (ref:183-191) by a fixed sequence: queue = host [n][queue_len], or NULL to go back
 * to the Philox stream. */
ST_API int st_host_set_piece_queue(StHostEnv *h, const uint8_t *queue, int32_t queue_len);
/* TetrisEnv.reset (ref:405-411) / TetrisEnv.step (ref:397-403) / engine.render() through _observation
 * (ref:317-321, 413-433) with HOST buffers, as a NumPy caller of the reference sees them.  obs/info/mask may be
 * NULL.  Copies (or maps, see st_host_set_zero_copy) the inputs in, launches, brings the results out, synchronises. */
ST_API int st_host_reset(StHostEnv *h, const uint8_t *mask, void *obs);
ST_API int st_host_step(StHostEnv *h, const uint8_t *actions, void *obs, float *reward, uint8_t *done, int32_t *info);
ST_API int st_host_observe(StHostEnv *h, int32_t draw_piece, void *obs);
/* Pipelined TetrisEnv.step (ref:397-403) for callers that keep the link busy (the loop of README.md:43-51 with the
 * next actions ready before the previous results are consumed, or two env groups stepped alternately):
 * st_host_step_async copies `actions` [n] (any host memory, free again on return), enqueues the step and ONE
 * device-to-host copy of its results into one of two page-locked slots owned by the handle, and returns without
 * waiting; at most two steps may be in flight.  st_host_wait blocks until the OLDEST step in flight is in host
 * memory and returns pointers into its slot: obs [n][st_obs_elems], reward [n], done [n], info [n][ST_INFO_WORDS]
 * (any out-pointer may be NULL).  The pointers stay valid until the second st_host_step_async after this wait.
 * The step kernel of call t+1 runs while the results of call t cross PCIe.  Synchronous st_host_* calls first
 * drain (and discard) whatever is still in flight. */
ST_API int st_host_step_async(StHostEnv *h, const uint8_t *actions);
ST_API int st_host_wait(StHostEnv *h, const void **obs, const float **reward, const uint8_t **done,
                        const int32_t **info);
/* When a host buffer passed to st_host_step is page-locked (cudaHostAlloc / cudaHostRegister / torch
 * pin_memory), the kernel can read / write it in place over PCIe instead of staging through device memory.
 * mask = OR of ST_ZC_*; pageable buffers always take the staging path.  Default: everything in place when one
 * step's outputs are at most 8 MiB, staging copies (copy engine at link rate) above that.  The pipelined calls
 * use the copy engine (measured on B200 at 4096 envs x 865 B: kernel stores into host memory reach 39 GB/s, one
 * cudaMemcpyAsync of the slot the link rate of ~55 GB/s); ST_ZC_PIPELINED selects in-place kernel writes there too. */
#define ST_ZC_ACTIONS 1 /* kernel reads the actions from host memory */
#define ST_ZC_SMALL 2   /* reward, done, info written straight to host memory */
#define ST_ZC_OBS 4     /* observations written straight to host memory */
#define ST_ZC_PIPELINED 8 /* st_host_step_async: the kernel writes the whole result slot in place */
ST_API int st_host_set_zero_copy(StHostEnv *h, int32_t mask);
/* Page-locked, device-mapped host memory for the buffers of st_host_step (NumPy callers have no pinned
 * allocator of their own); zero-filled.  NULL on failure. */
ST_API void *st_host_alloc_pinned(size_t bytes);
ST_API void st_host_free_pinned(void *ptr);
/* Re-key the piece stream (the `reset(seed=...)` of gym >= 0.26); effective from the next spawn. */
ST_API int st_host_set_seed(StHostEnv *h, uint64_t seed);
ST_API int st_host_render(StHostEnv *h, int32_t draw_piece, int32_t size, uint8_t *out);
ST_API int st_host_get_state(StHostEnv *h, uint8_t *boards, int32_t *scalars);
ST_API int st_host_set_state(StHostEnv *h, const uint8_t *boards, const int32_t *scalars);
/* Reads and clears the sticky device error flag; stats_out (u64[ST_STATS_WORDS]) may be NULL. */
ST_API int st_host_poll(StHostEnv *h, int32_t *error_flag_out, unsigned long long *stats_out);

/* ---- misc ------------------------------------------------------------------------- */
ST_API const char *st_last_error(void);
/* Name of the kernel a single-step launch of (cfg, n) runs: ram mode switches from the warp-per-env kernel to the
 * thread-per-env kernel for large batches (for reports and profiles). */
ST_API const char *st_step_kernel_name(const StConfig *cfg, int64_t n);
ST_API int st_abi_version(void);
/* Number of kernels this library has launched in this process (all threads). */
ST_API unsigned long long st_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SIMPLETETRIS_B200_H */
