// C ABI of include/simpletetris_b200.h: argument checking, geometry, launches, and the host-buffer handle.
#include "../../include/simpletetris_b200.h"
#include "st_internal.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, const char *detail = "")
{
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
int fail_cuda(cudaError_t e, const char *where)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return (int)e;
}

bool config_ok(const StConfig *c)
{
    return c && c->width >= 1 && c->width <= ST_MAX_WIDTH && c->height >= 1 && c->height <= ST_MAX_HEIGHT &&
           c->obs_type >= ST_OBS_RAM && c->obs_type <= ST_OBS_RGB && c->device >= 0;
}

int row_bytes(const StConfig *c) { return c->width <= 16 ? 2 : 4; }

// Fills everything of Params that derives from StConfig; buffers are left zero.
int make_params(const StConfig *c, const StAux *aux, int64_t n, st::Params *out)
{
    if (!config_ok(c)) return fail(ST_E_INVALID, "invalid StConfig (width 1..32, height 1..63, obs_type 0..2)%s");
    if (n < 0) return fail(ST_E_INVALID, "n < 0%s");
    st::Params p;
    memset(&p, 0, sizeof(p));
    p.W = c->width;
    p.H = c->height;
    p.lock_mod = (c->lock_delay > 0 ? c->lock_delay : 0) + 1;
    p.step_reset = c->step_reset != 0;
    p.auto_reset = c->auto_reset != 0;
    p.reward_step = c->reward_step != 0;
    p.pen_height = c->penalise_height != 0;
    p.pen_height_inc = c->penalise_height_increase != 0;
    p.adv_clears = c->advanced_clears != 0;
    p.high_scoring = c->high_scoring != 0;
    p.pen_holes = c->penalise_holes != 0;
    p.pen_holes_inc = c->penalise_holes_increase != 0;
    p.fullmask = c->width == 32 ? 0xffffffffu : ((1u << c->width) - 1u);
    p.row_bytes = row_bytes(c);
    p.stride = (int)st_state_stride(c);
    p.seed_lo = (uint32_t)c->seed;
    p.seed_hi = (uint32_t)(c->seed >> 32);
    p.env_id_base = c->env_id_base;
    p.obs_elems = (int)st_obs_elems(c);
    p.obs_u8 = c->obs_u8 != 0;
    // convert_grayscale geometry at size 84 (ref:84-94); the transposed array is (H, W)
    const int size = st::kImage;
    const int limiting = c->width > c->height ? c->width : c->height;
    const int gap = size / 100 + 1;
    const int bs = (size - 2 * gap) / limiting - gap;
    p.gap = gap;
    p.pitch = bs + gap;
    p.inner_v = gap + p.pitch * c->height;
    p.inner_h = gap + p.pitch * c->width;
    p.pad_top = (size - p.inner_v) / 2;
    p.pad_left = (size - p.inner_h) / 2;
    p.inv_h20 = ((1u << 20) + c->height - 1) / c->height;
    p.inv_sw20 = ((1u << 20) + (p.stride >> 2) - 1) / (p.stride >> 2);
    p.inv_hq20 = (c->height % 4 == 0) ? ((1u << 20) + c->height / 4 - 1) / (c->height / 4) : 0;
    if (aux) {
        p.queue = aux->piece_queue;
        p.queue_len = aux->queue_len;
        p.err = aux->error_flag;
        p.stats = aux->stats;
        p.term_obs = aux->terminal_obs;
        if (p.queue && p.queue_len <= 0) return fail(ST_E_INVALID, "piece_queue with queue_len <= 0%s");
    }
    p.n = n;
    p.T = 1;
    *out = p;
    return 0;
}

int use_device(const StConfig *c)
{
    static thread_local int current = -1;
    if (current != c->device) {
        cudaError_t e = cudaSetDevice(c->device);
        if (e != cudaSuccess) return fail_cuda(e, "cudaSetDevice");
        current = c->device;
    }
    return 0;
}

}  // namespace

extern "C" {

int64_t st_state_stride(const StConfig *cfg)
{
    if (!config_ok(cfg)) return -1;
    const int64_t b = 4 * ST_STATE_WORDS + (int64_t)cfg->height * row_bytes(cfg);
    return (b + 3) & ~(int64_t)3;
}

int64_t st_obs_elems(const StConfig *cfg)
{
    if (!config_ok(cfg)) return -1;
    if (cfg->obs_type == ST_OBS_RAM) return (int64_t)cfg->width * cfg->height;
    return cfg->obs_type == ST_OBS_GRAYSCALE ? 84 * 84 : 84 * 84 * 3;
}

int st_init(const StConfig *cfg, void *state, int64_t n, void *stream)
{
    st::Params p;
    if (int rc = make_params(cfg, nullptr, n, &p)) return rc;
    if (!state && n) return fail(ST_E_INVALID, "state is NULL%s");
    if (int rc = use_device(cfg)) return rc;
    p.state = (unsigned char *)state;
    cudaError_t e = st::launch_init(p, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "st_init");
}

int st_reset(const StConfig *cfg, void *state, const uint8_t *mask, void *obs, const StAux *aux, int64_t n,
             void *stream)
{
    st::Params p;
    if (int rc = make_params(cfg, aux, n, &p)) return rc;
    if (!state && n) return fail(ST_E_INVALID, "state is NULL%s");
    if (int rc = use_device(cfg)) return rc;
    p.state = (unsigned char *)state;
    p.mask = mask;
    p.obs = obs;
    p.mode = st::MODE_RESET;
    cudaError_t e = st::launch_main(p, cfg->obs_type, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "st_reset");
}

int st_step_many(const StConfig *cfg, void *state, const uint8_t *actions, int32_t T, void *obs,
                 int64_t obs_t_stride, float *reward, uint8_t *done, int32_t *info, int64_t info_t_stride,
                 const StAux *aux, int64_t n, void *stream)
{
    st::Params p;
    if (int rc = make_params(cfg, aux, n, &p)) return rc;
    if (n && (!state || !actions || !reward || !done)) return fail(ST_E_INVALID, "state/actions/reward/done is NULL%s");
    if (T < 1) return fail(ST_E_INVALID, "T < 1%s");
    if (((uintptr_t)state | (uintptr_t)obs) & 15) return fail(ST_E_INVALID, "state and obs must be 16-byte aligned%s");
    if (int rc = use_device(cfg)) return rc;
    p.state = (unsigned char *)state;
    p.actions = actions;
    p.obs = obs;
    p.reward = reward;
    p.done = done;
    p.info = info;
    p.T = T;
    p.obs_t_stride = obs_t_stride;
    p.info_t_stride = info_t_stride;
    p.mode = st::MODE_STEP;
    cudaError_t e = st::launch_main(p, cfg->obs_type, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "st_step");
}

int st_step(const StConfig *cfg, void *state, const uint8_t *actions, void *obs, float *reward, uint8_t *done,
            int32_t *info, const StAux *aux, int64_t n, void *stream)
{
    return st_step_many(cfg, state, actions, 1, obs, 0, reward, done, info, 0, aux, n, stream);
}

int st_observe(const StConfig *cfg, const void *state, int32_t draw_piece, void *obs, int64_t n, void *stream)
{
    st::Params p;
    if (int rc = make_params(cfg, nullptr, n, &p)) return rc;
    if (n && (!state || !obs)) return fail(ST_E_INVALID, "state/obs is NULL%s");
    if (int rc = use_device(cfg)) return rc;
    p.state = (unsigned char *)const_cast<void *>(state);
    p.obs = obs;
    p.mode = st::MODE_OBSERVE;
    p.draw_piece = draw_piece;
    cudaError_t e = st::launch_main(p, cfg->obs_type, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "st_observe");
}

int st_render(const StConfig *cfg, const void *state, int32_t draw_piece, int32_t size, uint8_t *out, int64_t n,
              void *stream)
{
    st::Params p;
    if (int rc = make_params(cfg, nullptr, n, &p)) return rc;
    if (n && (!state || !out)) return fail(ST_E_INVALID, "state/out is NULL%s");
    if (size < 4 || size > 4096) return fail(ST_E_INVALID, "render size outside 4..4096%s");
    if (n >= (1ll << 31)) return fail(ST_E_INVALID, "n too large%s");
    if (int rc = use_device(cfg)) return rc;
    p.state = (unsigned char *)const_cast<void *>(state);
    p.draw_piece = draw_piece;
    cudaError_t e = st::launch_render(p, size, out, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "st_render");
}

int st_get_state(const StConfig *cfg, const void *state, uint8_t *boards, int32_t *scalars, int64_t n, void *stream)
{
    st::Params p;
    if (int rc = make_params(cfg, nullptr, n, &p)) return rc;
    if (!state && n) return fail(ST_E_INVALID, "state is NULL%s");
    if (int rc = use_device(cfg)) return rc;
    p.state = (unsigned char *)const_cast<void *>(state);
    cudaError_t e = st::launch_get_state(p, boards, scalars, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "st_get_state");
}

int st_set_state(const StConfig *cfg, void *state, const uint8_t *boards, const int32_t *scalars, int64_t n,
                 void *stream)
{
    st::Params p;
    if (int rc = make_params(cfg, nullptr, n, &p)) return rc;
    if (!state && n) return fail(ST_E_INVALID, "state is NULL%s");
    if (int rc = use_device(cfg)) return rc;
    p.state = (unsigned char *)state;
    cudaError_t e = st::launch_set_state(p, boards, scalars, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "st_set_state");
}

// ---------------------------------------------------------------------------------------------
// Host-buffer handle
// ---------------------------------------------------------------------------------------------
struct StHostEnv {
    StConfig cfg;
    int64_t n;
    int64_t obs_elems;
    size_t obs_esz;  // 4 (float32, reference dtype) or 1 (uint8 mode)
    cudaStream_t stream;
    void *state;
    uint8_t *actions;
    void *obs;
    float *reward;
    uint8_t *done;
    int32_t *info;
    uint8_t *mask;
    uint8_t *queue;
    int32_t queue_len;
    int32_t *err;
    unsigned long long *stats;
    uint8_t *boards;
    int32_t *scalars;
    int zero_copy;  // ST_ZC_* mask
    // cache of the last host pointers seen and their device-side aliases (NULL = not page-locked)
    const void *zc_host[5];
    void *zc_dev[5];
};

#define HCHECK(call, where)                                  \
    do {                                                     \
        cudaError_t e__ = (call);                            \
        if (e__ != cudaSuccess) return fail_cuda(e__, where); \
    } while (0)

static StAux host_aux(const StHostEnv *h)
{
    StAux a;
    memset(&a, 0, sizeof(a));
    a.piece_queue = h->queue;
    a.queue_len = h->queue_len;
    a.error_flag = h->err;
    a.stats = h->stats;
    return a;
}

// Device-side alias of a page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) host pointer, or
// NULL for pageable memory.  One driver query per distinct pointer; steady-state calls hit the cache.
static void *mapped_alias(StHostEnv *h, int slot, const void *host)
{
    if (!host) return nullptr;
    if (h->zc_host[slot] == host) return h->zc_dev[slot];
    cudaPointerAttributes a;
    void *dev = nullptr;
    if (cudaPointerGetAttributes(&a, host) == cudaSuccess && a.type == cudaMemoryTypeHost) dev = a.devicePointer;
    else cudaGetLastError();
    h->zc_host[slot] = host;
    h->zc_dev[slot] = dev;
    return dev;
}

void *st_host_alloc_pinned(size_t bytes)
{
    void *ptr = nullptr;
    cudaError_t e = cudaHostAlloc(&ptr, bytes ? bytes : 1, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
        fail_cuda(e, "cudaHostAlloc");
        return nullptr;
    }
    memset(ptr, 0, bytes);
    return ptr;
}

void st_host_free_pinned(void *ptr)
{
    if (ptr) cudaFreeHost(ptr);
}

int st_host_set_seed(StHostEnv *h, uint64_t seed)
{
    if (!h) return fail(ST_E_INVALID, "NULL handle%s");
    h->cfg.seed = seed;  // takes effect at the next spawn: the stream is keyed by (seed, env id, piece index)
    return 0;
}

int st_host_set_zero_copy(StHostEnv *h, int32_t mask)
{
    if (!h) return fail(ST_E_INVALID, "NULL handle%s");
    h->zero_copy = mask;
    return 0;
}

StHostEnv *st_host_create(const StConfig *cfg, int64_t n)
{
    if (!config_ok(cfg) || n < 1) {
        fail(ST_E_INVALID, "st_host_create: invalid config or n < 1%s");
        return nullptr;
    }
    if (use_device(cfg)) return nullptr;
    StHostEnv *h = new (std::nothrow) StHostEnv();
    if (!h) return nullptr;
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->n = n;
    h->obs_elems = st_obs_elems(cfg);
    h->obs_esz = cfg->obs_u8 ? 1 : 4;
    // small batches: the kernel writes straight into page-locked caller buffers (saves four copy launches);
    // large ones: the copy engine moves the observation block at link rate
    h->zero_copy = (n * (h->obs_elems * (int64_t)h->obs_esz + 65) <= (8ll << 20)) ? (ST_ZC_ACTIONS | ST_ZC_SMALL | ST_ZC_OBS) : 0;
    bool ok = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) == cudaSuccess;
    ok = ok && cudaMalloc(&h->state, (size_t)(st_state_stride(cfg) * n)) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&h->actions, (size_t)n) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&h->obs, (size_t)(h->obs_elems * n) * h->obs_esz) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&h->reward, (size_t)n * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&h->done, (size_t)n) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&h->info, (size_t)n * ST_INFO_WORDS * sizeof(int32_t)) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&h->mask, (size_t)n) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&h->err, sizeof(int32_t)) == cudaSuccess;
    ok = ok && cudaMalloc((void **)&h->stats, ST_STATS_WORDS * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMemsetAsync(h->err, 0, sizeof(int32_t), h->stream) == cudaSuccess;
    ok = ok && cudaMemsetAsync(h->stats, 0, ST_STATS_WORDS * sizeof(unsigned long long), h->stream) == cudaSuccess;
    ok = ok && st_init(&h->cfg, h->state, n, h->stream) == 0;
    ok = ok && cudaStreamSynchronize(h->stream) == cudaSuccess;
    if (!ok) {
        if (!g_err[0]) fail_cuda(cudaGetLastError(), "st_host_create");
        st_host_destroy(h);
        return nullptr;
    }
    return h;
}

void st_host_destroy(StHostEnv *h)
{
    if (!h) return;
    use_device(&h->cfg);
    if (h->stream) cudaStreamSynchronize(h->stream);
    cudaFree(h->state); cudaFree(h->actions); cudaFree(h->obs); cudaFree(h->reward); cudaFree(h->done);
    cudaFree(h->info); cudaFree(h->mask); cudaFree(h->queue); cudaFree(h->err); cudaFree(h->stats);
    cudaFree(h->boards); cudaFree(h->scalars);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int st_host_set_piece_queue(StHostEnv *h, const uint8_t *queue, int32_t queue_len)
{
    if (!h) return fail(ST_E_INVALID, "NULL handle%s");
    if (int rc = use_device(&h->cfg)) return rc;
    HCHECK(cudaStreamSynchronize(h->stream), "sync");
    cudaFree(h->queue);
    h->queue = nullptr;
    h->queue_len = 0;
    if (!queue) return 0;
    if (queue_len <= 0) return fail(ST_E_INVALID, "queue_len <= 0%s");
    HCHECK(cudaMalloc((void **)&h->queue, (size_t)h->n * queue_len), "cudaMalloc(queue)");
    HCHECK(cudaMemcpyAsync(h->queue, queue, (size_t)h->n * queue_len, cudaMemcpyHostToDevice, h->stream), "H2D queue");
    HCHECK(cudaStreamSynchronize(h->stream), "sync");
    h->queue_len = queue_len;
    return 0;
}

int st_host_reset(StHostEnv *h, const uint8_t *mask, void *obs)
{
    if (!h) return fail(ST_E_INVALID, "NULL handle%s");
    if (int rc = use_device(&h->cfg)) return rc;
    if (mask) HCHECK(cudaMemcpyAsync(h->mask, mask, (size_t)h->n, cudaMemcpyHostToDevice, h->stream), "H2D mask");
    StAux aux = host_aux(h);
    if (int rc = st_reset(&h->cfg, h->state, mask ? h->mask : nullptr, h->obs, &aux, h->n, h->stream)) return rc;
    if (obs)
        HCHECK(cudaMemcpyAsync(obs, h->obs, (size_t)(h->obs_elems * h->n) * h->obs_esz, cudaMemcpyDeviceToHost,
                               h->stream), "D2H obs");
    HCHECK(cudaStreamSynchronize(h->stream), "st_host_reset");
    return 0;
}

int st_host_step(StHostEnv *h, const uint8_t *actions, void *obs, float *reward, uint8_t *done, int32_t *info)
{
    if (!h || !actions) return fail(ST_E_INVALID, "NULL handle/actions%s");
    if (int rc = use_device(&h->cfg)) return rc;
    // Page-locked caller buffers can be read / written by the kernel itself over PCIe (no staging copy, no
    // extra copy launches); pageable ones go through the handle's device buffers and cudaMemcpyAsync.
    const int zc = h->zero_copy;
    const uint8_t *d_act = (zc & ST_ZC_ACTIONS) ? (const uint8_t *)mapped_alias(h, 0, actions) : nullptr;
    void *d_obs = (zc & ST_ZC_OBS) ? mapped_alias(h, 1, obs) : nullptr;
    float *d_rew = (zc & ST_ZC_SMALL) ? (float *)mapped_alias(h, 2, reward) : nullptr;
    uint8_t *d_done = (zc & ST_ZC_SMALL) ? (uint8_t *)mapped_alias(h, 3, done) : nullptr;
    int32_t *d_info = (zc & ST_ZC_SMALL) ? (int32_t *)mapped_alias(h, 4, info) : nullptr;
    if (!d_act) {
        HCHECK(cudaMemcpyAsync(h->actions, actions, (size_t)h->n, cudaMemcpyHostToDevice, h->stream), "H2D actions");
        d_act = h->actions;
    }
    StAux aux = host_aux(h);
    if (int rc = st_step(&h->cfg, h->state, d_act, d_obs ? d_obs : h->obs, d_rew ? d_rew : h->reward,
                         d_done ? d_done : h->done, info ? (d_info ? d_info : h->info) : nullptr, &aux, h->n,
                         h->stream))
        return rc;
    if (obs && !d_obs)
        HCHECK(cudaMemcpyAsync(obs, h->obs, (size_t)(h->obs_elems * h->n) * h->obs_esz, cudaMemcpyDeviceToHost,
                               h->stream), "D2H obs");
    if (reward && !d_rew)
        HCHECK(cudaMemcpyAsync(reward, h->reward, (size_t)h->n * sizeof(float), cudaMemcpyDeviceToHost, h->stream),
               "D2H reward");
    if (done && !d_done) HCHECK(cudaMemcpyAsync(done, h->done, (size_t)h->n, cudaMemcpyDeviceToHost, h->stream), "D2H done");
    if (info && !d_info)
        HCHECK(cudaMemcpyAsync(info, h->info, (size_t)h->n * ST_INFO_WORDS * sizeof(int32_t), cudaMemcpyDeviceToHost,
                               h->stream), "D2H info");
    HCHECK(cudaStreamSynchronize(h->stream), "st_host_step");
    return 0;
}

int st_host_observe(StHostEnv *h, int32_t draw_piece, void *obs)
{
    if (!h || !obs) return fail(ST_E_INVALID, "NULL handle/obs%s");
    if (int rc = use_device(&h->cfg)) return rc;
    if (int rc = st_observe(&h->cfg, h->state, draw_piece, h->obs, h->n, h->stream)) return rc;
    HCHECK(cudaMemcpyAsync(obs, h->obs, (size_t)(h->obs_elems * h->n) * h->obs_esz, cudaMemcpyDeviceToHost,
                           h->stream), "D2H obs");
    HCHECK(cudaStreamSynchronize(h->stream), "st_host_observe");
    return 0;
}

int st_host_render(StHostEnv *h, int32_t draw_piece, int32_t size, uint8_t *out)
{
    if (!h || !out) return fail(ST_E_INVALID, "NULL handle/out%s");
    if (int rc = use_device(&h->cfg)) return rc;
    const size_t nb = (size_t)h->n * size * size * 3;
    uint8_t *d = nullptr;
    HCHECK(cudaMalloc((void **)&d, nb), "cudaMalloc(render)");
    int rc = st_render(&h->cfg, h->state, draw_piece, size, d, h->n, h->stream);
    cudaError_t e = rc ? cudaSuccess : cudaMemcpyAsync(out, d, nb, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e2 = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (rc) return rc;
    if (e != cudaSuccess) return fail_cuda(e, "D2H render");
    if (e2 != cudaSuccess) return fail_cuda(e2, "st_host_render");
    return 0;
}

static int host_scratch(StHostEnv *h)
{
    const size_t nb = (size_t)h->n * h->cfg.width * h->cfg.height;
    if (!h->boards) HCHECK(cudaMalloc((void **)&h->boards, nb), "cudaMalloc(boards)");
    if (!h->scalars) HCHECK(cudaMalloc((void **)&h->scalars, (size_t)h->n * ST_UNPACKED_WORDS * 4), "cudaMalloc(scalars)");
    return 0;
}

int st_host_get_state(StHostEnv *h, uint8_t *boards, int32_t *scalars)
{
    if (!h) return fail(ST_E_INVALID, "NULL handle%s");
    if (int rc = use_device(&h->cfg)) return rc;
    if (int rc = host_scratch(h)) return rc;
    if (int rc = st_get_state(&h->cfg, h->state, boards ? h->boards : nullptr, scalars ? h->scalars : nullptr, h->n,
                              h->stream))
        return rc;
    const size_t nb = (size_t)h->n * h->cfg.width * h->cfg.height;
    if (boards) HCHECK(cudaMemcpyAsync(boards, h->boards, nb, cudaMemcpyDeviceToHost, h->stream), "D2H boards");
    if (scalars)
        HCHECK(cudaMemcpyAsync(scalars, h->scalars, (size_t)h->n * ST_UNPACKED_WORDS * 4, cudaMemcpyDeviceToHost,
                               h->stream), "D2H scalars");
    HCHECK(cudaStreamSynchronize(h->stream), "st_host_get_state");
    return 0;
}

int st_host_set_state(StHostEnv *h, const uint8_t *boards, const int32_t *scalars)
{
    if (!h) return fail(ST_E_INVALID, "NULL handle%s");
    if (int rc = use_device(&h->cfg)) return rc;
    if (int rc = host_scratch(h)) return rc;
    const size_t nb = (size_t)h->n * h->cfg.width * h->cfg.height;
    if (boards) HCHECK(cudaMemcpyAsync(h->boards, boards, nb, cudaMemcpyHostToDevice, h->stream), "H2D boards");
    if (scalars)
        HCHECK(cudaMemcpyAsync(h->scalars, scalars, (size_t)h->n * ST_UNPACKED_WORDS * 4, cudaMemcpyHostToDevice,
                               h->stream), "H2D scalars");
    if (int rc = st_set_state(&h->cfg, h->state, boards ? h->boards : nullptr, scalars ? h->scalars : nullptr, h->n,
                              h->stream))
        return rc;
    HCHECK(cudaStreamSynchronize(h->stream), "st_host_set_state");
    return 0;
}

int st_host_poll(StHostEnv *h, int32_t *error_flag_out, unsigned long long *stats_out)
{
    if (!h) return fail(ST_E_INVALID, "NULL handle%s");
    if (int rc = use_device(&h->cfg)) return rc;
    int32_t e = 0;
    HCHECK(cudaMemcpyAsync(&e, h->err, sizeof(e), cudaMemcpyDeviceToHost, h->stream), "D2H err");
    if (stats_out)
        HCHECK(cudaMemcpyAsync(stats_out, h->stats, ST_STATS_WORDS * sizeof(unsigned long long),
                               cudaMemcpyDeviceToHost, h->stream), "D2H stats");
    HCHECK(cudaMemsetAsync(h->err, 0, sizeof(int32_t), h->stream), "memset err");
    HCHECK(cudaStreamSynchronize(h->stream), "st_host_poll");
    if (error_flag_out) *error_flag_out = e;
    return 0;
}

const char *st_step_kernel_name(const StConfig *cfg, int64_t n)
{
    st::Params p;
    if (make_params(cfg, nullptr, n, &p)) return "";
    return st::step_kernel_name(p, cfg->obs_type);
}

const char *st_last_error(void) { return g_err; }
int st_abi_version(void) { return ST_ABI_VERSION; }
unsigned long long st_launch_count(void) { return st::launch_count(); }

}  // extern "C"
