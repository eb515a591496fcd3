/*
 * st_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU restatement of the reference's algorithm for the SimpleTetris
 * step path (reference: gym_simpletetris/envs/tetris_env.py, cited below as
 * `ref:LINE`).  It deliberately keeps the reference's own data model — a dense
 * (W,H) float64 board indexed [x][y], pieces as four (i,j) offsets, rotation by
 * coordinate swap, np.repeat/np.insert image construction — so that it checks
 * the CUDA bitboard kernels through a different formulation.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker / CPU baseline.
 * The product (gym_simpletetris_b200) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py runs this file
 * against the unmodified reference executed in the build container, and
 * tests/test_oracle_golden.py checks it against the .npz fixtures under tests/golden/, which
 * tests/golden/make_golden.py generated from the unmodified reference.
 *
 * One deliberate extension (north star, not reference): the piece source.  The
 * reference draws from Python's global Mersenne Twister (ref:187); here the
 * uniform draw comes either from an injected piece queue (parity runs) or from
 * a Philox4x32-10 counter-based stream keyed by (seed, global env id) and
 * indexed by the lifetime piece count — the same stream the CUDA kernel uses.
 * The weighting rule itself (ref:183-191) is restated exactly.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- piece tables (ref:10-19) ------------------------------------------- */
static const int SHAPES[7][4][2] = {
    /* T */ {{0, 0}, {-1, 0}, {1, 0}, {0, -1}},
    /* J */ {{0, 0}, {-1, 0}, {0, -1}, {0, -2}},
    /* L */ {{0, 0}, {1, 0}, {0, -1}, {0, -2}},
    /* Z */ {{0, 0}, {-1, 0}, {0, -1}, {1, -1}},
    /* S */ {{0, 0}, {-1, -1}, {0, -1}, {1, 0}},
    /* I */ {{0, 0}, {0, -1}, {0, -2}, {0, -3}},
    /* O */ {{0, 0}, {0, -1}, {-1, 0}, {-1, -1}},
};

typedef struct {
    int c[4][2];
} Shape;

typedef struct OrEnv {
    int W, H;
    int lock_delay, step_reset;
    int reward_step, penalise_height, penalise_height_increase;
    int advanced_clears, high_scoring, penalise_holes, penalise_holes_increase;
    double *board; /* [x*H + y], ref:140 */
    Shape shape;
    int shape_id; /* -1 = None (ref:170-172) */
    int ax, ay;   /* anchor (ref:196; W/2 truncated as ref:244 does) */
    int time, score, holes, lines_cleared, piece_height, n_deaths; /* ref:165-173 */
    int ld;        /* _lock_delay, ref:176 */
    int counts[7]; /* shape_counts, ref:181 */
    /* piece source */
    uint64_t seed;
    int64_t env_id;
    const uint8_t *queue;
    int qlen;
    int error; /* sticky: 1 = queue exhausted, 2 = bad action, 4 = step with no piece */
} OrEnv;

/* ---- Philox4x32-10 (Salmon et al. 2011), same constants as the kernel ---- */
static uint32_t philox_draw(uint64_t seed, int64_t env_id, uint32_t index)
{
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)(uint64_t)env_id, c1 = (uint32_t)((uint64_t)env_id >> 32);
    uint32_t c2 = index, c3 = 0;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c0;
}

/* ---- rotated (ref:22-26) -------------------------------------------------- */
static Shape rotated(Shape s, int cclk)
{
    Shape o;
    for (int k = 0; k < 4; ++k) {
        int i = s.c[k][0], j = s.c[k][1];
        if (cclk) { o.c[k][0] = -j; o.c[k][1] = i; }
        else      { o.c[k][0] = j;  o.c[k][1] = -i; }
    }
    return o;
}

/* ---- is_occupied (ref:29-36) ---------------------------------------------- */
static int is_occupied(const OrEnv *e, Shape s, int ax, int ay)
{
    for (int k = 0; k < 4; ++k) {
        int x = ax + s.c[k][0], y = ay + s.c[k][1];
        if (y < 0) continue;
        if (x < 0 || x >= e->W || y >= e->H || e->board[x * e->H + y] != 0.0) return 1;
    }
    return 0;
}

/* ---- the seven actions (ref:39-73) ---------------------------------------- */
static void act_left(const OrEnv *e, Shape *s, int *ax, int *ay)
{ if (!is_occupied(e, *s, *ax - 1, *ay)) *ax -= 1; }
static void act_right(const OrEnv *e, Shape *s, int *ax, int *ay)
{ if (!is_occupied(e, *s, *ax + 1, *ay)) *ax += 1; }
static void act_soft_drop(const OrEnv *e, Shape *s, int *ax, int *ay)
{ if (!is_occupied(e, *s, *ax, *ay + 1)) *ay += 1; }
static void act_hard_drop(const OrEnv *e, Shape *s, int *ax, int *ay)
{
    for (;;) { /* ref:54-59 */
        int y0 = *ay;
        act_soft_drop(e, s, ax, ay);
        if (*ay == y0) return;
    }
}
static void act_rotate_left(const OrEnv *e, Shape *s, int *ax, int *ay)
{ Shape n = rotated(*s, 0); if (!is_occupied(e, n, *ax, *ay)) *s = n; }
static void act_rotate_right(const OrEnv *e, Shape *s, int *ax, int *ay)
{ Shape n = rotated(*s, 1); if (!is_occupied(e, n, *ax, *ay)) *s = n; }

/* ---- _choose_shape (ref:183-191) ------------------------------------------ */
static int choose_shape(OrEnv *e)
{
    int64_t total = 0;
    int maxm = e->counts[0];
    for (int i = 0; i < 7; ++i) { total += e->counts[i]; if (e->counts[i] > maxm) maxm = e->counts[i]; }
    if (e->queue) {
        if (total >= e->qlen) { e->error |= 1; return e->queue[total % e->qlen] % 7; }
        return e->queue[total] % 7;
    }
    int m[7];
    int64_t S = 0;
    for (int i = 0; i < 7; ++i) { m[i] = 5 + maxm - e->counts[i]; S += m[i]; }
    /* r = random.randint(1, S): one uniform draw per piece (ref:187) */
    uint32_t u = philox_draw(e->seed, e->env_id, (uint32_t)total);
    int64_t r = 1 + (int64_t)(((uint64_t)u * (uint64_t)S) >> 32);
    for (int i = 0; i < 7; ++i) { r -= m[i]; if (r <= 0) return i; }
    return 6; /* unreachable */
}

/* ---- _new_piece (ref:193-200) --------------------------------------------- */
static void new_piece(OrEnv *e)
{
    e->ax = e->W / 2; /* float W/2, truncated by int() at ref:244 before any use */
    e->ay = 0;
    e->shape_id = choose_shape(e);
    e->counts[e->shape_id] += 1;
    memcpy(&e->shape, SHAPES[e->shape_id], sizeof(Shape));
}

/* ---- _set_piece (ref:323-327) --------------------------------------------- */
static void set_piece(OrEnv *e, int on)
{
    for (int k = 0; k < 4; ++k) {
        int x = e->shape.c[k][0] + e->ax, y = e->shape.c[k][1] + e->ay;
        if (x < e->W && x >= 0 && y < e->H && y >= 0) e->board[x * e->H + y] = on ? 1.0 : 0.0;
    }
}

/* ---- _clear_lines (ref:205-216) ------------------------------------------- */
static int clear_lines(OrEnv *e)
{
    int W = e->W, H = e->H, n = 0;
    int *can_clear = (int *)malloc(sizeof(int) * H);
    double *nb = (double *)calloc((size_t)W * H, sizeof(double));
    for (int i = 0; i < H; ++i) {
        int all = 1;
        for (int x = 0; x < W; ++x) if (e->board[x * H + i] == 0.0) { all = 0; break; }
        can_clear[i] = all;
        n += all;
    }
    int j = H - 1;
    for (int i = H - 1; i >= 0; --i) {
        if (!can_clear[i]) {
            for (int x = 0; x < W; ++x) nb[x * H + j] = e->board[x * H + i];
            j -= 1;
        }
    }
    e->lines_cleared += n;
    memcpy(e->board, nb, sizeof(double) * W * H);
    free(nb);
    free(can_clear);
    return n;
}

/* ---- _count_holes (ref:218-220): cumsum along y, times ~board -------------- */
static int count_holes(OrEnv *e)
{
    int n = 0;
    for (int x = 0; x < e->W; ++x) {
        double cs = 0.0;
        for (int y = 0; y < e->H; ++y) {
            double b = e->board[x * e->H + y];
            cs += b;
            if (cs * (b != 0.0 ? 0.0 : 1.0) != 0.0) n += 1;
        }
    }
    e->holes = n;
    return n;
}

/* sum(np.any(board, axis=0)) (ref:287,289): number of rows y with any cell */
static int nonempty_rows(const OrEnv *e)
{
    int n = 0;
    for (int y = 0; y < e->H; ++y) {
        int any = 0;
        for (int x = 0; x < e->W; ++x) if (e->board[x * e->H + y] != 0.0) { any = 1; break; }
        n += any;
    }
    return n;
}

/* ---- TetrisEngine.step (ref:243-304); state_out = board with piece drawn ---- */
static int engine_step(OrEnv *e, int action, double *state_out, double *reward_out, int *done_out)
{
    if (e->shape_id < 0) { e->error |= 4; *reward_out = 0; *done_out = 0; return -1; }
    switch (action) { /* ref:245, map ref:152-160 */
    case 0: act_left(e, &e->shape, &e->ax, &e->ay); break;
    case 1: act_right(e, &e->shape, &e->ax, &e->ay); break;
    case 2: act_hard_drop(e, &e->shape, &e->ax, &e->ay); break;
    case 3: act_soft_drop(e, &e->shape, &e->ax, &e->ay); break;
    case 4: act_rotate_left(e, &e->shape, &e->ax, &e->ay); break;
    case 5: act_rotate_right(e, &e->shape, &e->ax, &e->ay); break;
    case 6: break;
    default: e->error |= 2; break; /* reference raises KeyError; batched contract: idle + flag */
    }
    int nay = e->ay; /* ref:247 gravity */
    if (!is_occupied(e, e->shape, e->ax, e->ay + 1)) nay = e->ay + 1;
    if (e->step_reset && nay != e->ay) e->ld = 0; /* ref:248-249 */
    e->ay = nay;

    e->time += 1; /* ref:253 */
    double reward = e->reward_step ? 1 : 0; /* ref:256 */
    int done = 0;
    if (is_occupied(e, e->shape, e->ax, e->ay + 1)) { /* ref:259 */
        int L = (e->lock_delay > 0 ? e->lock_delay : 0) + 1;
        e->ld = (e->ld + 1) % L; /* ref:175,260 */
        if (e->ld == 0) { /* ref:262 */
            set_piece(e, 1);
            int k = clear_lines(e);
            static const int scores[5] = {0, 40, 100, 300, 1200};
            if (e->advanced_clears) { /* ref:266-275 */
                int kk = k > 4 ? 4 : k; /* reference would IndexError for k>4; unreachable in play */
                reward += 2.5 * scores[kk];
                e->score += scores[kk];
            } else if (e->high_scoring) {
                reward += 1000 * k;
                e->score += k;
            } else {
                reward += 100 * k;
                e->score += k;
            }
            int top = 0; /* np.any(board[:,0]) ref:277 */
            for (int x = 0; x < e->W; ++x) if (e->board[x * e->H + 0] != 0.0) top = 1;
            if (top) {
                count_holes(e);
                e->n_deaths += 1;
                done = 1;
                reward = -100;
            } else {
                int old_holes = e->holes;
                count_holes(e);
                if (e->penalise_height) { /* ref:286-292 */
                    reward -= nonempty_rows(e);
                } else if (e->penalise_height_increase) {
                    int nh = nonempty_rows(e);
                    if (nh > e->piece_height) reward -= 10 * (nh - e->piece_height);
                    e->piece_height = nh;
                }
                if (e->penalise_holes) reward -= 5 * e->holes; /* ref:294-297 */
                else if (e->penalise_holes_increase) reward -= 5 * (e->holes - old_holes);
                new_piece(e); /* ref:299 */
            }
        }
    }
    set_piece(e, 1); /* ref:301-303 */
    if (state_out) memcpy(state_out, e->board, sizeof(double) * e->W * e->H);
    set_piece(e, 0);
    *reward_out = reward;
    *done_out = done;
    return 0;
}

/* ---- clear (ref:306-315) --------------------------------------------------- */
static void engine_clear(OrEnv *e)
{
    e->time = 0; e->score = 0; e->holes = 0; e->lines_cleared = 0; e->piece_height = 0;
    new_piece(e);
    memset(e->board, 0, sizeof(double) * e->W * e->H);
}

/* ---- numpy helpers used by convert_grayscale ------------------------------- */
typedef struct { int r, c; uint8_t *d; } Arr;
static Arr arr_new(int r, int c) { Arr a = {r, c, (uint8_t *)calloc((size_t)(r > 0 ? r : 1) * (c > 0 ? c : 1), 1)}; return a; }
/* np.repeat(a, k, axis) */
static Arr arr_repeat(Arr a, int k, int axis)
{
    Arr o = axis == 0 ? arr_new(a.r * k, a.c) : arr_new(a.r, a.c * k);
    for (int i = 0; i < o.r; ++i)
        for (int j = 0; j < o.c; ++j)
            o.d[i * o.c + j] = axis == 0 ? a.d[(i / k) * a.c + j] : a.d[i * a.c + j / k];
    free(a.d);
    return o;
}
/* np.insert(a, idx (sorted, may repeat), value, axis): rows/cols go in BEFORE original index idx[t] */
static Arr arr_insert(Arr a, const int *idx, int nidx, uint8_t value, int axis)
{
    int n = axis == 0 ? a.r : a.c;
    Arr o = axis == 0 ? arr_new(a.r + nidx, a.c) : arr_new(a.r, a.c + nidx);
    int *src = (int *)malloc(sizeof(int) * (n + nidx + 1)); /* -1 = inserted */
    int pos = 0, t = 0;
    for (int p = 0; p <= n; ++p) {
        while (t < nidx && idx[t] == p) { src[pos++] = -1; ++t; }
        if (p < n) src[pos++] = p;
    }
    for (int i = 0; i < o.r; ++i)
        for (int j = 0; j < o.c; ++j) {
            int s = axis == 0 ? src[i] : src[j];
            o.d[i * o.c + j] = s < 0 ? value : (axis == 0 ? a.d[s * a.c + j] : a.d[i * a.c + s]);
        }
    free(src);
    free(a.d);
    return o;
}

/* ---- convert_grayscale (ref:76-114); out = size*size uint8, row-major ------- */
static void convert_grayscale(const double *board, int W, int H, int size, uint8_t *out)
{
    const uint8_t border = 0, background = 128, piece = 190;
    Arr a = arr_new(H, W); /* np.transpose(np.array(board, uint8)) ref:81-82 */
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) a.d[y * W + x] = (uint8_t)board[x * H + y];
    int s0 = H, s1 = W;
    int limiting = s0 > s1 ? s0 : s1;
    int gap = size / 100 + 1;
    int bs = (size - 2 * gap) / limiting - gap;
    int inner_w = gap + (bs + gap) * s0;
    int inner_h = gap + (bs + gap) * s1;
    int pad_w = (size - inner_w) / 2;
    int pad_h = (size - inner_h) / 2;
    for (int i = 0; i < H * W; ++i) { /* ref:96-97 (sequential, as in the reference) */
        if (a.d[i] == 0) a.d[i] = background;
        if (a.d[i] == 1) a.d[i] = piece;
    }
    a = arr_repeat(a, bs, 0);
    a = arr_repeat(a, bs, 1);
    int n0 = (s0 + 1) * gap, n1 = (s1 + 1) * gap;
    int *idx = (int *)malloc(sizeof(int) * (size_t)((n0 > n1 ? n0 : n1) + 2 * size + 4));
    int t = 0;
    for (int x = 0; x <= s0; ++x) for (int g = 0; g < gap; ++g) idx[t++] = bs * x;
    a = arr_insert(a, idx, t, background, 0); /* ref:102-104 */
    t = 0;
    for (int x = 0; x <= s1; ++x) for (int g = 0; g < gap; ++g) idx[t++] = bs * x;
    a = arr_insert(a, idx, t, background, 1); /* ref:105-107 */
    int len = a.r; /* ref:109 */
    t = 0;
    for (int g = 0; g < pad_w; ++g) idx[t++] = 0;
    for (int g = 0; g < size - (pad_w + len); ++g) idx[t++] = len;
    a = arr_insert(a, idx, t, border, 0);
    len = a.c; /* ref:111 */
    t = 0;
    for (int g = 0; g < pad_h; ++g) idx[t++] = 0;
    for (int g = 0; g < size - (pad_h + len); ++g) idx[t++] = len;
    a = arr_insert(a, idx, t, border, 1);
    memcpy(out, a.d, (size_t)size * size);
    free(idx);
    free(a.d);
}

/* ---- TetrisEnv._observation + float32 cast (ref:400,413-433) ---------------- */
/* obs_type: 0 ram, 1 grayscale, 2 rgb.  extend_dims only adds a trailing 1. */
static void observation(const OrEnv *e, const double *state, int obs_type, float *out)
{
    if (obs_type == 0) {
        for (int i = 0; i < e->W * e->H; ++i) out[i] = (float)state[i];
        return;
    }
    uint8_t *img = (uint8_t *)malloc(84 * 84);
    convert_grayscale(state, e->W, e->H, 84, img);
    if (obs_type == 1) {
        for (int i = 0; i < 84 * 84; ++i) out[i] = (float)img[i];
    } else { /* convert_grayscale_rgb (ref:117-122): np.repeat(gray[...,None], 3, axis=2) */
        for (int i = 0; i < 84 * 84; ++i) { out[3 * i] = out[3 * i + 1] = out[3 * i + 2] = (float)img[i]; }
    }
    free(img);
}

/* ======================= exported API (ctypes) ============================== */
#define API __attribute__((visibility("default")))

API int or_obs_elems(int W, int H, int obs_type)
{ return obs_type == 0 ? W * H : (obs_type == 1 ? 84 * 84 : 84 * 84 * 3); }

/* flags: [reward_step, penalise_height, penalise_height_increase, advanced_clears,
 *         high_scoring, penalise_holes, penalise_holes_increase] */
API OrEnv *or_create(int W, int H, int lock_delay, int step_reset, const int *flags, uint64_t seed, int64_t env_id)
{
    OrEnv *e = (OrEnv *)calloc(1, sizeof(OrEnv));
    e->W = W; e->H = H; e->lock_delay = lock_delay; e->step_reset = step_reset;
    e->reward_step = flags[0]; e->penalise_height = flags[1]; e->penalise_height_increase = flags[2];
    e->advanced_clears = flags[3]; e->high_scoring = flags[4]; e->penalise_holes = flags[5];
    e->penalise_holes_increase = flags[6];
    e->board = (double *)calloc((size_t)W * H, sizeof(double));
    e->time = -1; e->score = -1; /* ref:165-166 */
    e->shape_id = -1;
    e->seed = seed; e->env_id = env_id;
    return e;
}
API void or_destroy(OrEnv *e) { if (e) { free(e->board); free(e); } }
API void or_set_queue(OrEnv *e, const uint8_t *queue, int qlen) { e->queue = queue; e->qlen = qlen; }
API int or_error(const OrEnv *e) { return e->error; }

/* reset(): obs of the EMPTY board, piece not drawn (ref:405-408, 313-315) */
API void or_reset(OrEnv *e, int obs_type, float *obs)
{
    engine_clear(e);
    if (obs) observation(e, e->board, obs_type, obs);
}

API int or_step(OrEnv *e, int action, int obs_type, float *obs, double *reward, int *done)
{
    double *state = (double *)malloc(sizeof(double) * e->W * e->H);
    int rc = engine_step(e, action, state, reward, done);
    if (rc == 0 && obs) observation(e, state, obs_type, obs);
    free(state);
    return rc;
}

/* info (ref:232-241) as 13 ints: time, piece id, score, lines, holes, deaths, counts[7] */
API void or_info(const OrEnv *e, int *out)
{
    out[0] = e->time; out[1] = e->shape_id; out[2] = e->score; out[3] = e->lines_cleared;
    out[4] = e->holes; out[5] = e->n_deaths;
    for (int i = 0; i < 7; ++i) out[6 + i] = e->counts[i];
}

/* debug state: piece (id, rot, x, y), ld, piece_height; rot = number of rotate_left applications */
API void or_get_piece(const OrEnv *e, int *out)
{
    out[0] = e->shape_id; out[1] = -1; out[2] = e->ax; out[3] = e->ay; out[4] = e->ld; out[5] = e->piece_height;
    if (e->shape_id >= 0) {
        Shape s; memcpy(&s, SHAPES[e->shape_id], sizeof(Shape));
        for (int r = 0; r < 4; ++r) {
            if (memcmp(&s, &e->shape, sizeof(Shape)) == 0) { out[1] = r; break; }
            s = rotated(s, 0);
        }
    }
}
API void or_set_piece(OrEnv *e, int id, int rot, int x, int y)
{
    e->shape_id = id;
    memcpy(&e->shape, SHAPES[id], sizeof(Shape));
    for (int r = 0; r < (rot & 3); ++r) e->shape = rotated(e->shape, 0);
    e->ax = x; e->ay = y;
}
/* TetrisEngine.render (ref:317-321): _set_piece(True); copy; _set_piece(False).  -1 when there is no piece. */
API int or_render(OrEnv *e, double *out)
{
    if (e->shape_id < 0) return -1;
    set_piece(e, 1);
    memcpy(out, e->board, sizeof(double) * e->W * e->H);
    set_piece(e, 0);
    return 0;
}
API void or_get_board(const OrEnv *e, double *out) { memcpy(out, e->board, sizeof(double) * e->W * e->H); }
API void or_set_board(OrEnv *e, const double *in) { memcpy(e->board, in, sizeof(double) * e->W * e->H); }
/* counters: time, score, lines, holes, piece_height, deaths, ld, counts[7] (14 ints) */
API void or_get_counters(const OrEnv *e, int *o)
{
    o[0] = e->time; o[1] = e->score; o[2] = e->lines_cleared; o[3] = e->holes; o[4] = e->piece_height;
    o[5] = e->n_deaths; o[6] = e->ld;
    for (int i = 0; i < 7; ++i) o[7 + i] = e->counts[i];
}
API void or_set_counters(OrEnv *e, const int *o)
{
    e->time = o[0]; e->score = o[1]; e->lines_cleared = o[2]; e->holes = o[3]; e->piece_height = o[4];
    e->n_deaths = o[5]; e->ld = o[6];
    for (int i = 0; i < 7; ++i) e->counts[i] = o[7 + i];
}
API void or_convert_grayscale(const double *board, int W, int H, int size, uint8_t *out)
{ convert_grayscale(board, W, H, size, out); }

/*
 * Vector rollout with gym<=0.25 auto-reset semantics (what the product's VecEnv
 * does on the GPU): n envs with global ids env_id_base.., T steps, actions
 * [T][n] uint8.  Per step and env writes obs [n][elems] (overwritten each step,
 * final step's obs remains), reward [T][n] f32, done [T][n] u8 and, if non-null,
 * info [T][n][13] i32 taken BEFORE the auto-reset.  `envs` may be NULL (fresh
 * envs are created, reset, and destroyed) or an array of n live envs.
 * Threads: OpenMP over envs (nthreads<=0 -> default).  Returns sticky error OR.
 */
API int or_rollout(OrEnv **envs, int W, int H, int lock_delay, int step_reset, const int *flags,
                   int obs_type, uint64_t seed, int64_t env_id_base, int64_t n, int T,
                   const uint8_t *actions, float *obs, float *reward, uint8_t *done, int32_t *info,
                   int auto_reset, int nthreads)
{
    int err = 0;
    int elems = or_obs_elems(W, H, obs_type);
#ifdef _OPENMP
    omp_set_num_threads(nthreads > 0 ? nthreads : omp_get_num_procs());
#endif
#pragma omp parallel for schedule(static) reduction(| : err)
    for (int64_t i = 0; i < n; ++i) {
        OrEnv *e = envs ? envs[i] : or_create(W, H, lock_delay, step_reset, flags, seed, env_id_base + i);
        float *o = obs ? obs + (size_t)i * elems : NULL;
        if (!envs) or_reset(e, obs_type, o);
        for (int t = 0; t < T; ++t) {
            double r; int d;
            or_step(e, actions[(size_t)t * n + i], obs_type, o, &r, &d);
            if (reward) reward[(size_t)t * n + i] = (float)r;
            if (done) done[(size_t)t * n + i] = (uint8_t)d;
            if (info) or_info(e, info + ((size_t)t * n + i) * 13);
            if (d && auto_reset) or_reset(e, obs_type, o);
        }
        err |= e->error;
        if (!envs) or_destroy(e);
    }
    return err;
}

API int or_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
