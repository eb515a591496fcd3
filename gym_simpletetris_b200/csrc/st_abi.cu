// C ABI of include/simpletetris_b200.h: argument checking, geometry, launches, and the host-buffer handle.
#include "../../include/simpletetris_b200.h"
#include "st_internal.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

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

// 32-bit words per board column of the env record: bit y of column x = cell (x, y)
int col_words(const StConfig *c) { return c->height <= 31 ? 1 : 2; }

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
    p.col_words = col_words(c);
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
    p.inv_nq32 = (c->height % 4 == 0) ? (uint32_t)(((1ull << 32) + c->width * c->height / 4 - 1) / (c->width * c->height / 4)) : 0;
    p.inv_w20 = ((1u << 20) + c->width - 1) / c->width;
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

// Every entry point runs on cfg->device and leaves the caller's current device as it found it (torch or user code
// may change the current device between two calls, so nothing about it is cached across calls).
struct DeviceGuard {
    int prev = -1, rc = 0;
    bool switched = false;
    explicit DeviceGuard(int want)
    {
        cudaError_t e = cudaGetDevice(&prev);
        if (e == cudaSuccess && prev != want) {
            e = cudaSetDevice(want);
            switched = e == cudaSuccess;
        }
        if (e != cudaSuccess) rc = fail_cuda(e, "cudaSetDevice");
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};
#define USE_DEVICE(cfgptr)                 \
    DeviceGuard dev_guard__((cfgptr)->device); \
    if (dev_guard__.rc) return dev_guard__.rc

}  // namespace

extern "C" {

int64_t st_state_stride(const StConfig *cfg)
{
    if (!config_ok(cfg)) return -1;
    return 4 * ST_STATE_WORDS + (int64_t)cfg->width * 4 * col_words(cfg);
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
    USE_DEVICE(cfg);
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
    USE_DEVICE(cfg);
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
    if (T > 1) {  // every step's observation block must keep the alignment the vector / bulk stores rely on:
        // 16 bytes for images (TMA bulk stores), four elements for ram boards with H % 4 == 0 (one store per four
        // cells), one element otherwise (scalar stores)
        const int64_t esz = cfg->obs_u8 ? 1 : 4;
        const int64_t align = cfg->obs_type != ST_OBS_RAM ? 16 : (cfg->height % 4 == 0 ? 4 * esz : esz);
        if (obs && obs_t_stride != 0 && (obs_t_stride < p.obs_elems * n || (obs_t_stride * esz) % align != 0))
            return fail(ST_E_INVALID, "obs_t_stride must be 0 or >= n*st_obs_elems, and keep every step's block aligned "
                                      "(16 bytes for images, 4 elements for ram boards with height % 4 == 0)%s");
        if (info && info_t_stride != 0 && (info_t_stride < n * ST_INFO_WORDS || info_t_stride % ST_INFO_WORDS != 0))
            return fail(ST_E_INVALID, "info_t_stride must be 0 or a multiple of ST_INFO_WORDS >= n*ST_INFO_WORDS%s");
        if (aux && aux->terminal_obs)
            return fail(ST_E_INVALID, "terminal_obs is a one-step buffer: not supported with T > 1%s");
    }
    USE_DEVICE(cfg);
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
    USE_DEVICE(cfg);
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
    USE_DEVICE(cfg);
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
    USE_DEVICE(cfg);
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
    USE_DEVICE(cfg);
    p.state = (unsigned char *)state;
    cudaError_t e = st::launch_set_state(p, boards, scalars, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : fail_cuda(e, "st_set_state");
}

// ---------------------------------------------------------------------------------------------
// Host-buffer handle
// ---------------------------------------------------------------------------------------------
// One in-flight step of the pipelined form (st_host_step_async / st_host_wait): a device block and its page-locked
// host mirror with the same layout  obs | info | reward | done  (one cudaMemcpyAsync brings a step's results out).
struct StHostSlot {
    unsigned char *dev;   // device block
    unsigned char *host;  // page-locked mirror (on the GPU's NUMA node when the kernel allows it)
    void *host_alias;     // device-side alias of `host` (zero-copy writes), NULL if not mapped
    uint8_t *act_host;    // page-locked copy of the caller's actions for this step
    uint8_t *act_dev;
    cudaEvent_t kernel_done, copied;
    int host_is_mmap;
};

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
    uint8_t *render_buf;   // grow-only device buffer of st_host_render
    size_t render_bytes;
    int zero_copy;  // ST_ZC_* mask
    // pipelined step: two slots, a copy stream beside the compute stream
    cudaStream_t copy_stream;
    StHostSlot slot[2];
    size_t off_info, off_reward, off_done, slot_bytes;
    int slots_ready;
    unsigned long long head, tail;  // steps enqueued / steps waited for
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

// Device-side alias of a page-locked (cudaHostAlloc / cudaHostRegister / torch pin_memory) host pointer, or NULL
// for pageable memory.  Queried on every call: a caller may free a pinned buffer and get a pageable one at the
// same address, so nothing is remembered across calls (one driver query per pointer, ~0.2 us).
static void *mapped_alias(const void *host)
{
    if (!host) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) == cudaSuccess && a.type == cudaMemoryTypeHost) return a.devicePointer;
    cudaGetLastError();
    return nullptr;
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

// NUMA node of a CUDA device (sysfs), -1 if unknown.
static int device_numa_node(int device)
{
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    for (char *c = bus; *c; ++c)
        if (*c >= 'A' && *c <= 'Z') *c = (char)(*c - 'A' + 'a');
    char path[96];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", bus);
    FILE *f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

// Page-locked host block for a pipeline slot.  Eight processes writing 3.5 MB per step each into page-locked memory
// that all sits on one NUMA node is what limited the 8-GPU end-to-end rate in round 1, so the pages are bound to the
// NUMA node the GPU hangs off (mmap + mbind + cudaHostRegister); plain cudaHostAlloc when that is not possible.
static unsigned char *slot_host_alloc(size_t bytes, int device, int *is_mmap)
{
    *is_mmap = 0;
    static const bool numa_off = getenv("ST_B200_NO_NUMA") != nullptr;
    const int node = numa_off ? -1 : device_numa_node(device);
    if (node >= 0 && node < 1024) {
        const size_t len = (bytes + 4095) & ~(size_t)4095;
        void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            unsigned long maskbits[16] = {0};
            maskbits[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
            // MPOL_PREFERRED = 1: falls back to other nodes instead of failing when the node is full
            syscall(SYS_mbind, p, len, 1, maskbits, (unsigned long)(8 * sizeof(maskbits)), 0u);
            memset(p, 0, len);
            if (cudaHostRegister(p, len, cudaHostRegisterPortable | cudaHostRegisterMapped) == cudaSuccess) {
                *is_mmap = 1;
                return (unsigned char *)p;
            }
            cudaGetLastError();
            munmap(p, len);
        }
    }
    return (unsigned char *)st_host_alloc_pinned(bytes);
}

static void slot_host_free(unsigned char *p, size_t bytes, int is_mmap)
{
    if (!p) return;
    if (is_mmap) {
        cudaHostUnregister(p);
        munmap(p, (bytes + 4095) & ~(size_t)4095);
    } else {
        cudaFreeHost(p);
    }
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
    if (h->head != h->tail) return fail(ST_E_INVALID, "st_host_set_zero_copy with steps in flight%s");
    h->zero_copy = mask;
    return 0;
}

StHostEnv *st_host_create(const StConfig *cfg, int64_t n)
{
    if (!config_ok(cfg) || n < 1) {
        fail(ST_E_INVALID, "st_host_create: invalid config or n < 1%s");
        return nullptr;
    }
    DeviceGuard guard(cfg->device);
    if (guard.rc) return nullptr;
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

static void free_slots(StHostEnv *h)
{
    for (int s = 0; s < 2; ++s) {
        StHostSlot &sl = h->slot[s];
        cudaFree(sl.dev);
        cudaFree(sl.act_dev);
        slot_host_free(sl.host, h->slot_bytes, sl.host_is_mmap);
        if (sl.act_host) cudaFreeHost(sl.act_host);
        if (sl.kernel_done) cudaEventDestroy(sl.kernel_done);
        if (sl.copied) cudaEventDestroy(sl.copied);
        memset(&sl, 0, sizeof(sl));
    }
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    h->copy_stream = nullptr;
    h->slots_ready = 0;
}

void st_host_destroy(StHostEnv *h)
{
    if (!h) return;
    DeviceGuard guard(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);
    free_slots(h);
    cudaFree(h->state); cudaFree(h->actions); cudaFree(h->obs); cudaFree(h->reward); cudaFree(h->done);
    cudaFree(h->info); cudaFree(h->mask); cudaFree(h->queue); cudaFree(h->err); cudaFree(h->stats);
    cudaFree(h->boards); cudaFree(h->scalars); cudaFree(h->render_buf);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// Steps enqueued by st_host_step_async that nobody has waited for yet must not race with a synchronous call.
static int drain_pipeline(StHostEnv *h)
{
    if (h->head == h->tail) return 0;
    HCHECK(cudaStreamSynchronize(h->stream), "drain");
    HCHECK(cudaStreamSynchronize(h->copy_stream), "drain");
    h->tail = h->head;
    return 0;
}

int st_host_set_piece_queue(StHostEnv *h, const uint8_t *queue, int32_t queue_len)
{
    if (!h) return fail(ST_E_INVALID, "NULL handle%s");
    USE_DEVICE(&h->cfg);
    if (int rc = drain_pipeline(h)) return rc;
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
    USE_DEVICE(&h->cfg);
    if (int rc = drain_pipeline(h)) return rc;
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
    USE_DEVICE(&h->cfg);
    if (int rc = drain_pipeline(h)) return rc;
    // Page-locked caller buffers can be read / written by the kernel itself over PCIe (no staging copy, no
    // extra copy launches); pageable ones go through the handle's device buffers and cudaMemcpyAsync.
    const int zc = h->zero_copy;
    const uint8_t *d_act = (zc & ST_ZC_ACTIONS) ? (const uint8_t *)mapped_alias(actions) : nullptr;
    void *d_obs = (zc & ST_ZC_OBS) ? mapped_alias(obs) : nullptr;
    float *d_rew = (zc & ST_ZC_SMALL) ? (float *)mapped_alias(reward) : nullptr;
    uint8_t *d_done = (zc & ST_ZC_SMALL) ? (uint8_t *)mapped_alias(done) : nullptr;
    int32_t *d_info = (zc & ST_ZC_SMALL) ? (int32_t *)mapped_alias(info) : nullptr;
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

// ---- pipelined step ---------------------------------------------------------------------------------------
static int ensure_slots(StHostEnv *h)
{
    if (h->slots_ready) return 0;
    const size_t n = (size_t)h->n;
    const size_t obs_bytes = ((size_t)h->obs_elems * n * h->obs_esz + 15) & ~(size_t)15;
    h->off_info = obs_bytes;
    h->off_reward = h->off_info + n * ST_INFO_WORDS * sizeof(int32_t);
    h->off_done = h->off_reward + n * sizeof(float);
    h->slot_bytes = (h->off_done + n + 15) & ~(size_t)15;
    HCHECK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking), "cudaStreamCreate");
    for (int s = 0; s < 2; ++s) {
        StHostSlot &sl = h->slot[s];
        HCHECK(cudaMalloc((void **)&sl.dev, h->slot_bytes), "cudaMalloc(slot)");
        HCHECK(cudaMalloc((void **)&sl.act_dev, n), "cudaMalloc(slot actions)");
        sl.host = slot_host_alloc(h->slot_bytes, h->cfg.device, &sl.host_is_mmap);
        sl.act_host = (uint8_t *)st_host_alloc_pinned(n);
        if (!sl.host || !sl.act_host) return fail(ST_E_INVALID, "page-locked slot allocation failed%s");
        sl.host_alias = mapped_alias(sl.host);
        HCHECK(cudaEventCreateWithFlags(&sl.kernel_done, cudaEventDisableTiming), "cudaEventCreate");
        HCHECK(cudaEventCreateWithFlags(&sl.copied, cudaEventDisableTiming), "cudaEventCreate");
    }
    h->slots_ready = 1;
    return 0;
}

int st_host_step_async(StHostEnv *h, const uint8_t *actions)
{
    if (!h || !actions) return fail(ST_E_INVALID, "NULL handle/actions%s");
    if (h->head - h->tail >= 2) return fail(ST_E_INVALID, "st_host_step_async: two steps already in flight, call st_host_wait%s");
    USE_DEVICE(&h->cfg);
    if (int rc = ensure_slots(h)) {
        free_slots(h);
        return rc;
    }
    StHostSlot &sl = h->slot[h->head & 1];
    const size_t n = (size_t)h->n;
    memcpy(sl.act_host, actions, n);  // the caller's array is free again when this returns
    // the slot's device block is still being read by the D2H copy of the step before last
    if (h->head >= 2) HCHECK(cudaStreamWaitEvent(h->stream, sl.copied, 0), "cudaStreamWaitEvent");
    HCHECK(cudaMemcpyAsync(sl.act_dev, sl.act_host, n, cudaMemcpyHostToDevice, h->stream), "H2D actions");
    const bool zc = (h->zero_copy & ST_ZC_PIPELINED) && sl.host_alias;
    unsigned char *blk = zc ? (unsigned char *)sl.host_alias : sl.dev;
    StAux aux = host_aux(h);
    if (int rc = st_step(&h->cfg, h->state, sl.act_dev, blk, (float *)(blk + h->off_reward), blk + h->off_done,
                         (int32_t *)(blk + h->off_info), &aux, h->n, h->stream))
        return rc;
    HCHECK(cudaEventRecord(sl.kernel_done, h->stream), "cudaEventRecord");
    if (!zc) {  // one copy brings obs | info | reward | done out while the next step's kernel runs
        HCHECK(cudaStreamWaitEvent(h->copy_stream, sl.kernel_done, 0), "cudaStreamWaitEvent");
        HCHECK(cudaMemcpyAsync(sl.host, sl.dev, h->off_done + n, cudaMemcpyDeviceToHost, h->copy_stream), "D2H slot");
        HCHECK(cudaEventRecord(sl.copied, h->copy_stream), "cudaEventRecord");
    } else {
        HCHECK(cudaEventRecord(sl.copied, h->stream), "cudaEventRecord");
    }
    h->head += 1;
    return 0;
}

int st_host_wait(StHostEnv *h, const void **obs, const float **reward, const uint8_t **done, const int32_t **info)
{
    if (!h) return fail(ST_E_INVALID, "NULL handle%s");
    if (h->head == h->tail) return fail(ST_E_INVALID, "st_host_wait: no step in flight%s");
    USE_DEVICE(&h->cfg);
    StHostSlot &sl = h->slot[h->tail & 1];
    HCHECK(cudaEventSynchronize(sl.copied), "st_host_wait");
    if (obs) *obs = sl.host;
    if (info) *info = (const int32_t *)(sl.host + h->off_info);
    if (reward) *reward = (const float *)(sl.host + h->off_reward);
    if (done) *done = sl.host + h->off_done;
    h->tail += 1;
    return 0;
}

int st_host_observe(StHostEnv *h, int32_t draw_piece, void *obs)
{
    if (!h || !obs) return fail(ST_E_INVALID, "NULL handle/obs%s");
    USE_DEVICE(&h->cfg);
    if (int rc = drain_pipeline(h)) return rc;
    if (int rc = st_observe(&h->cfg, h->state, draw_piece, h->obs, h->n, h->stream)) return rc;
    HCHECK(cudaMemcpyAsync(obs, h->obs, (size_t)(h->obs_elems * h->n) * h->obs_esz, cudaMemcpyDeviceToHost,
                           h->stream), "D2H obs");
    HCHECK(cudaStreamSynchronize(h->stream), "st_host_observe");
    return 0;
}

int st_host_render(StHostEnv *h, int32_t draw_piece, int32_t size, uint8_t *out)
{
    if (!h || !out) return fail(ST_E_INVALID, "NULL handle/out%s");
    if (size < 4 || size > 4096) return fail(ST_E_INVALID, "render size outside 4..4096%s");
    USE_DEVICE(&h->cfg);
    if (int rc = drain_pipeline(h)) return rc;
    const size_t nb = (size_t)h->n * size * size * 3;
    if (nb > h->render_bytes) {  // grow-only scratch: video logging calls this every frame
        HCHECK(cudaStreamSynchronize(h->stream), "sync");
        cudaFree(h->render_buf);
        h->render_buf = nullptr;
        h->render_bytes = 0;
        HCHECK(cudaMalloc((void **)&h->render_buf, nb), "cudaMalloc(render)");
        h->render_bytes = nb;
    }
    if (int rc = st_render(&h->cfg, h->state, draw_piece, size, h->render_buf, h->n, h->stream)) return rc;
    HCHECK(cudaMemcpyAsync(out, h->render_buf, nb, cudaMemcpyDeviceToHost, h->stream), "D2H render");
    HCHECK(cudaStreamSynchronize(h->stream), "st_host_render");
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
    USE_DEVICE(&h->cfg);
    if (int rc = drain_pipeline(h)) return rc;
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
    USE_DEVICE(&h->cfg);
    if (int rc = drain_pipeline(h)) return rc;
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
    USE_DEVICE(&h->cfg);
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
