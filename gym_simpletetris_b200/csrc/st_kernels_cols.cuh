// K1c — warp-per-env ram step on COLUMN LANES (included by st_kernels.cu after st_kernels_tpe.cuh).
//
// The env record keeps the board as one word per column (bit y of column x = cell (x, y)), and the ram observation is
// column-major too (ref:140, 421-424).  Here lane L of the env's warp owns board column L - 4; lanes 0..3 and the lanes
// from W + 4 up are WALL columns (all ones), so a piece cell left or right of the board finds a full column and the
// walls need no code of their own (ref:34).  The engine is the thread-per-env one (st_kernels_tpe.cuh) spread over
// the lanes:
//   * collisions for every anchor height at once: the lane whose column holds cell k of the piece contributes its
//     column moved by that cell's row offset, one REDUX.OR joins the four contributions (ref:29-36);
//   * lock: each lane ORs its share of the piece into its column; full rows = REDUX.AND, rows with any cell = REDUX.OR,
//     holes = REDUX.ADD of (H - top - popc) per lane; a cleared row is squeezed out of every column in parallel;
//   * the scalar part of the record (15 words) sits one word per lane, as in the row kernel (K1).
//   * scalars are broadcast with REDUX, not SHFL: a REDUX result is a uniform register, so the compiler knows every
//     branch on it is warp-uniform (a shuffle result is "divergent" to it and every `if` gets a convergence barrier).
// Control flow is uniform per warp (one env), so the lock / spawn / reset branches cost nothing when they are not
// taken.  Measured (ncu, 4096 envs of 10x20): 270 warp-instructions per env-step inside a T-step launch, 400 in a
// one-step launch, against 425 / 580 for K1, which pays a column<->row transpose at both ends of every step since the
// record went column-major; same reward table, Philox stream and error flags.
#pragma once

namespace st {

constexpr int kColsLaneOff = 4;   // lane of board column 0: a candidate anchor at x = -1 with a cell at i = -3 is lane 0
constexpr int kColsMaxW = 24;     // ... and x = W with i = +3 is lane W + 7 <= 31
#ifndef ST_COLS_WARPS
#define ST_COLS_WARPS 4  // warps (= envs) per CTA: tuning knob
#endif
constexpr int kColsWarpsPerCta = ST_COLS_WARPS;

// The four cells of a piece, one byte per field (a uniform index here — one piece per warp — so the constant cache
// broadcasts it, and a byte leaves a word with one PRMT):  i3 = byte k: column offset + 3;  s3 = byte k: row offset + 3;
// maxj3 = largest row offset + 3.
//   pat_lo / pat_hi = the piece by COLUMN: a 7-bit pattern per column offset d = i + 3 (bit j + 3 = cell (i, j)),
//   offsets 0..3 in pat_lo, 4..6 in pat_hi, 7 bits each.
struct ColsTab {
    uint32_t i3[28], s3[28], maxj3[28], pat_lo[28], pat_hi[28];
};
constexpr ColsTab make_cols_tab()
{
    const CellTab t = make_cell_tab();
    ColsTab c{};
    for (int s = 0; s < 28; ++s) {
        const uint32_t lo = (uint32_t)t.e[s];
        for (int k = 0; k < 4; ++k) {
            c.i3[s] |= ((lo >> (8 * k)) & 7u) << (8 * k);
            c.s3[s] |= ((lo >> (8 * k + 3)) & 7u) << (8 * k);
            const uint32_t d = (lo >> (8 * k)) & 7u, j3 = (lo >> (8 * k + 3)) & 7u;
            if (d < 4) c.pat_lo[s] |= 1u << (7 * d + j3);
            else c.pat_hi[s] |= 1u << (7 * (d - 4) + j3);
        }
        c.maxj3[s] = (uint32_t)(t.e[s] >> 32) & 15u;
    }
    return c;
}
__constant__ ColsTab c_cols = make_cols_tab();

struct ColsCells {
    int i3[4], s3[4], maxj;  // column offset + 3, row offset + 3 of the four cells; largest row offset
    uint32_t pat_lo, pat_hi;
};

__device__ __forceinline__ ColsCells cols_cells(int id, int rot)
{
    const uint32_t wi = c_cols.i3[id * 4 + rot], ws = c_cols.s3[id * 4 + rot];
    ColsCells c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c.i3[k] = (int)__byte_perm(wi, 0, 0x4440 + k);
        c.s3[k] = (int)__byte_perm(ws, 0, 0x4440 + k);
    }
    c.maxj = (int)c_cols.maxj3[id * 4 + rot] - 3;
    c.pat_lo = c_cols.pat_lo[id * 4 + rot];
    c.pat_hi = c_cols.pat_hi[id * 4 + rot];
    return c;
}

template <typename ColT> __device__ __forceinline__ ColT warp_or(ColT v);
template <> __device__ __forceinline__ uint32_t warp_or<uint32_t>(uint32_t v) { return __reduce_or_sync(FULL, v); }
template <> __device__ __forceinline__ unsigned long long warp_or<unsigned long long>(unsigned long long v)
{
    return ((unsigned long long)__reduce_or_sync(FULL, (uint32_t)(v >> 32)) << 32) | __reduce_or_sync(FULL, (uint32_t)v);
}
template <typename ColT> __device__ __forceinline__ ColT warp_and(ColT v);
template <> __device__ __forceinline__ uint32_t warp_and<uint32_t>(uint32_t v) { return __reduce_and_sync(FULL, v); }
template <> __device__ __forceinline__ unsigned long long warp_and<unsigned long long>(unsigned long long v)
{
    return ((unsigned long long)__reduce_and_sync(FULL, (uint32_t)(v >> 32)) << 32) | __reduce_and_sync(FULL, (uint32_t)v);
}
template <typename ColT> __device__ __forceinline__ ColT warp_get(ColT v, int src);
template <> __device__ __forceinline__ uint32_t warp_get<uint32_t>(uint32_t v, int src) { return __shfl_sync(FULL, v, src); }
template <> __device__ __forceinline__ unsigned long long warp_get<unsigned long long>(unsigned long long v, int src)
{
    return ((unsigned long long)__shfl_sync(FULL, (uint32_t)(v >> 32), src) << 32) | __shfl_sync(FULL, (uint32_t)v, src);
}

// Word `idx` of the per-lane scalar state, as a value the compiler KNOWS to be warp-uniform: REDUX writes a uniform
// register, so everything computed from it runs on the uniform datapath and every branch on it is a plain uniform
// branch (a shuffle result is "divergent" to the compiler, which then wraps each `if` in a convergence barrier).
__device__ __forceinline__ int uget(int sw, int lane, int idx) { return (int)__reduce_or_sync(FULL, lane == idx ? (uint32_t)sw : 0u); }

// _new_piece / _choose_shape (ref:183-200): spawn_piece of the row kernel with uniform reads
__device__ __forceinline__ int cols_spawn(int &sw, int lane, const Params &p, int e, int &errbits)
{
    const bool cnt = lane >= 8 && lane <= 14;
    const int total = __reduce_add_sync(FULL, cnt ? sw : 0);
    const int mx = __reduce_max_sync(FULL, cnt ? sw : 0);
    int id;
    if (p.queue) {
        int k = total;
        if (k >= p.queue_len) { errbits |= 1; k %= p.queue_len; }
        id = p.queue[(size_t)e * (unsigned)p.queue_len + k] % 7;
        id = (int)__reduce_or_sync(FULL, (uint32_t)id);
    } else {
        const int S = 35 + 7 * mx - total;  // sum of m_i = 5 + max - c_i (ref:186)
        const uint32_t u = philox_draw(p.seed_lo, p.seed_hi, (unsigned long long)(p.env_id_base + e), (uint32_t)total);
        const int r = 1 + (int)__umulhi(u, (uint32_t)S);  // uniform on [1, S] (ref:187)
        // smallest i with cumsum(m)[i] >= r (ref:188-191): lane 8 + i holds m_i, an inclusive scan over lanes 8..13
        int acc = (lane >= 8 && lane <= 13) ? 5 + mx - sw : 0;
#pragma unroll
        for (int d = 1; d < 8; d <<= 1) {
            const int up = __shfl_up_sync(FULL, acc, d);
            if (lane >= 8 + d) acc += up;
        }
        id = __popc(__ballot_sync(FULL, lane >= 8 && lane <= 13 && r > acc));
    }
    if (lane == 8 + id) sw += 1;
    return id;
}

// Bit y' set <=> is_occupied(shape, (x, y'), board) (ref:29-36), for every anchor height y' at once.  X3 = this lane's
// board column + 3 (wall lanes hold all ones).
template <typename ColT>
__device__ __forceinline__ ColT cols_collisions(ColT col, const ColsCells &c, int x, int X3, int W, int H)
{
    ColT mine = 0;
    const int d = X3 - x;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (d == c.i3[k]) mine |= ColOps<ColT>::by_rows(col, c.s3[k]);  // cells above the board fall off the shift (ref:32-33)
    if ((unsigned)(x + 1) > (unsigned)(W + 1)) {  // only an injected anchor: columns beyond the lanes are outside the board too
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((unsigned)(x + c.i3[k] - 3 + kColsLaneOff) > 31u) mine |= ColOps<ColT>::by_rows(~(ColT)0, c.s3[k]);
    }
    int fl = H - c.maxj;  // anchors whose lowest cell is at or below the floor (ref:34)
    fl = fl < 0 ? 0 : fl;
    return warp_or<ColT>(mine) | (~(ColT)0 << fl);
}

// This lane's share of the piece at anchor (x, y): its in-board cells in column X3 - 3 (ref:325-326 `0 <= y < height`):
// the column pattern of offset d = X3 - x, moved to row y.
template <typename ColT>
__device__ __forceinline__ ColT cols_piece(const ColsCells &c, int x, int y, int X3, int H)
{
    const int d = X3 - x;
    const uint32_t w = d < 4 ? c.pat_lo : c.pat_hi;
    uint32_t pat = (w >> (7 * (d & 3))) & 127u;
    if ((unsigned)d > 6u) pat = 0;
    constexpr int kYMax = sizeof(ColT) == 4 ? 56 : 66;  // an injected anchor far below the floor: every cell drops out
    y = y > kYMax ? kYMax : y;
    ColT m;
    if constexpr (sizeof(ColT) == 4) m = (ColT)(((unsigned long long)pat << y) >> 3);
    else m = y >= 3 ? ((ColT)pat << (y - 3)) : ((ColT)pat >> (3 - y));
    return m & (((ColT)1 << H) - 1);
}

// rows every board column has / rows any board column has / _count_holes (ref:218-220)
template <typename ColT>
__device__ __forceinline__ void cols_scan(ColT col, bool on_board, int H, ColT &full, ColT &any, int &holes)
{
    using Ops = ColOps<ColT>;
    const ColT hbit = (ColT)1 << H;
    full = warp_and<ColT>(on_board ? col : ~(ColT)0) & (hbit - 1);
    any = warp_or<ColT>(on_board ? col : (ColT)0);
    holes = __reduce_add_sync(FULL, on_board ? H + 1 - Ops::ffs(col | hbit) - Ops::popc(col) : 0);
}

// The lock branch (ref:262-299), uniform over the warp.
template <typename ColT>
__device__ __forceinline__ void cols_lock(ColT &col, int &sw, Piece &pc, const ColsCells &cl, int lane, int X3, bool on_board,
                                          const Params &p, int e, int &reward, int &done, int &errbits)
{
    using Ops = ColOps<ColT>;
    const int H = p.H;
    col |= cols_piece<ColT>(cl, pc.x, pc.y, X3, H);  // _set_piece(True) (ref:263); wall lanes are all ones already
    ColT full, any;
    int holes;
    cols_scan<ColT>(col, on_board, H, full, any, holes);
    const int k = Ops::popc(full);
    if (k) {  // _clear_lines (ref:205-216): top-most full row first, every column in parallel
        if (on_board) {
            ColT f = full;
            while (f) {
                const ColT bit = f & (~f + 1);
                f ^= bit;
                col = (col & ~(bit | (bit - 1))) | ((col & (bit - 1)) << 1);
            }
        }
        ColT dummy;
        cols_scan<ColT>(col, on_board, H, dummy, any, holes);
        if (lane == 4) sw += k;  // lines_cleared (ref:213)
    }
    const int nonempty = Ops::popc(any);
    int dscore;
    if (p.adv_clears) {  // ref:266-275
        const int kk = k > 4 ? 4 : k;
        dscore = kk == 0 ? 0 : kk == 1 ? 40 : kk == 2 ? 100 : kk == 3 ? 300 : 1200;
        reward += (dscore * 5) / 2;
    } else if (p.high_scoring) {
        dscore = k;
        reward += 1000 * k;
    } else {
        dscore = k;
        reward += 100 * k;
    }
    if (lane == 3) sw += dscore;
    const int old_holes = uget(sw, lane, 5);
    put(sw, lane, 5, holes);
    if (any & 1) {  // np.any(board[:, 0]) (ref:277-281)
        if (lane == 7) sw += 1;
        done = 1;
        reward = -100;
    } else {
        if (p.pen_height) {
            reward -= nonempty;
        } else if (p.pen_height_inc) {
            const int ph = uget(sw, lane, 6);
            if (nonempty > ph) reward -= 10 * (nonempty - ph);
            put(sw, lane, 6, nonempty);
        }
        if (p.pen_holes) reward -= 5 * holes;
        else if (p.pen_holes_inc) reward -= 5 * (holes - old_holes);
        pc.id = cols_spawn(sw, lane, p, e, errbits);  // ref:299
        pc.rot = 0; pc.x = p.W / 2; pc.y = 0;
    }
}

// ram observation of one env from the column lanes: float32 / uint8 [W][H] (ref:421-424, 400)
template <typename ColT>
__device__ __forceinline__ void cols_write_ram(ColT shown, void *out, const float4 *s_lut, const Params &p, int lane)
{
    const int H = p.H, W = p.W;
    const bool u8 = p.obs_u8 != 0;
    if ((H & 3) == 0) {
        const int hq = H >> 2, nq = W * hq;
        for (int q0 = 0; q0 < nq; q0 += 32) {  // uniform trip count: every lane takes part in the shuffles
            const int q = q0 + lane;
            const int x = (int)(((uint32_t)min(q, nq - 1) * p.inv_hq20) >> 20);
            const int k = min(q, nq - 1) - x * hq;
            const ColT c = warp_get<ColT>(shown, x + kColsLaneOff);
            const uint32_t nib = (uint32_t)(c >> (4 * k)) & 15u;
            if (q < nq) {
                if (u8) reinterpret_cast<uint32_t *>(out)[q] = (nib * 0x00204081u) & 0x01010101u;
                else reinterpret_cast<float4 *>(out)[q] = s_lut[nib];
            }
        }
    } else {
        const int nel = W * H;
        for (int i0 = 0; i0 < nel; i0 += 32) {
            const int i = i0 + lane;
            const int x = (int)(((uint32_t)min(i, nel - 1) * p.inv_h20) >> 20);
            const int y = min(i, nel - 1) - x * H;
            const ColT c = warp_get<ColT>(shown, x + kColsLaneOff);
            const bool on = ((c >> y) & 1) != 0;
            if (i < nel) {
                if (u8) reinterpret_cast<unsigned char *>(out)[i] = on ? 1 : 0;
                else reinterpret_cast<float *>(out)[i] = on ? 1.0f : 0.0f;
            }
        }
    }
}

// MINB = the min-blocks hint of __launch_bounds__, i.e. the register budget: the kernel is latency-bound, so the build
// with the most registers (most loads and table reads hoisted) is the fastest AS LONG AS the whole batch is resident in
// one wave; launch_cols picks the instantiation by occupancy (measured on B200, 10x20, us per launch, 72 / 63 / 53
// registers: 2048 envs 3.26 / 3.73 / 4.19, 4096 envs 3.89 / 4.28 / 4.55, 5120 envs 5.06 / 4.98 / 4.46).
template <typename ColT, int MINB>
__global__ void __launch_bounds__(32 * kColsWarpsPerCta, MINB) st_step_cols_kernel(const __grid_constant__ Params p)
{
    __shared__ __align__(16) float4 s_lut[16];
    constexpr int CW = ColOps<ColT>::kWords;
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31;
    const int warp = (int)__reduce_or_sync(FULL, threadIdx.x >> 5);  // uniform, and known to be
    if (threadIdx.x < 16)
        s_lut[threadIdx.x] = make_float4((threadIdx.x & 1) ? 1.0f : 0.0f, (threadIdx.x & 2) ? 1.0f : 0.0f,
                                         (threadIdx.x & 4) ? 1.0f : 0.0f, (threadIdx.x & 8) ? 1.0f : 0.0f);
    const int n = (int)p.n;
    const int e = (int)blockIdx.x * kColsWarpsPerCta + warp;
    const int W = p.W, H = p.H;
    const int X = lane - kColsLaneOff;
    const int X3 = X + 3;
    const bool on_board = (unsigned)X < (unsigned)W;
    __syncthreads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (e >= n) return;

    unsigned int action_u;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(action_u) : "l"(p.actions + e));
    uint32_t *rec_w = reinterpret_cast<uint32_t *>(p.state + (size_t)e * (unsigned)p.stride);
    int sw = lane < kStateWords ? (int)rec_w[lane] : 0;
    ColT col = ~(ColT)0;
    if (on_board) {
        if constexpr (CW == 1) col = rec_w[kStateWords + X];
        else col = (ColT)rec_w[kStateWords + 2 * X] | ((ColT)rec_w[kStateWords + 2 * X + 1] << 32);
    }
    const bool u8 = p.obs_u8 != 0;
    const size_t esz = u8 ? 1 : 4;
    // running output pointers (advance per step in st_step_many; an absent buffer stays unused behind its flag)
    const bool has_info = p.info != nullptr, has_obs = p.obs != nullptr, want_term = p.term_obs != nullptr;
    const uint8_t *act_p = p.actions + e;
    float *rew_p = p.reward + e;
    uint8_t *done_p = p.done + e;
    int32_t *info_p = p.info + (size_t)e * kStateWords + lane;
    char *obs_p = reinterpret_cast<char *>(p.obs) + (size_t)e * (unsigned)p.obs_elems * esz;
    char *term_p = reinterpret_cast<char *>(p.term_obs) + (size_t)e * (unsigned)p.obs_elems * esz;
    const long long obs_step = p.obs_t_stride * (long long)esz;
    // float32 observations of boards with at most 64 float4 slots (10x20: 50): each lane's one or two slots — which
    // column lane to read, which nibble of it — are the same every step
    const int hq = H >> 2, nq = W * hq;
    const bool fast_obs = has_obs && !u8 && (H & 3) == 0 && nq <= 64;
    int src0 = 0, sh0 = 0, src1 = 0, sh1 = 0;
    if (fast_obs) {
        const int q0 = min(lane, nq - 1), q1 = min(lane + 32, nq - 1);
        const int x0 = (int)(((uint32_t)q0 * p.inv_hq20) >> 20), x1 = (int)(((uint32_t)q1 * p.inv_hq20) >> 20);
        src0 = x0 + kColsLaneOff; sh0 = 4 * (q0 - x0 * hq);
        src1 = x1 + kColsLaneOff; sh1 = 4 * (q1 - x1 * hq);
    }
    int action = (int)__reduce_or_sync(FULL, action_u);
    int errbits = 0;
    Piece pc = unpack_piece(uget(sw, lane, 0));  // the piece and the lock-delay counter live in uniform registers
    int ld = uget(sw, lane, 1);

    for (int t = 0; t < p.T; ++t) {
        unsigned int next_u = 6u;  // next step's action: its miss overlaps this step
        if (t + 1 < p.T) asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(next_u) : "l"(act_p + n));
        // ---- TetrisEngine.step (ref:243-304) ----
        int reward = p.reward_step, done = 0;
        ColT pm = 0;  // this lane's share of the piece the returned state shows (ref:301)
        if (pc.id >= 7) {  // no piece yet: the reference would fail on shape None (ref:170-172,245)
            errbits |= 4;
        } else {
            if (action > 6) { errbits |= 2; action = 6; }
            int r2 = pc.rot, x2 = pc.x;
            if (action == 0) x2 -= 1;
            if (action == 1) x2 += 1;
            if (action == 4) r2 = (r2 + 1) & 3;
            if (action == 5) r2 = (r2 + 3) & 3;
            ColsCells cl = cols_cells(pc.id, r2);
            ColT cm = cols_collisions<ColT>(col, cl, x2, X3, W, H);
            const bool moved = (r2 != pc.rot) || (x2 != pc.x);
            if (moved && ((cm >> pc.y) & 1)) {  // blocked: stay (ref:41,46,64,69)
                cl = cols_cells(pc.id, pc.rot);
                cm = cols_collisions<ColT>(col, cl, pc.x, X3, W, H);
            } else {
                pc.rot = r2; pc.x = x2;
            }
            int y = pc.y;
            if (action == 3 && !((cm >> (y + 1)) & 1)) y += 1;  // soft_drop (ref:49-51)
            if (action == 2) {                                   // hard_drop (ref:54-59): first blocked height below
                const ColT below = cm >> (y + 1);
                if (below) y += ColOps<ColT>::ffs(below) - 1;
            }
            if (!((cm >> (y + 1)) & 1)) {                        // gravity (ref:247-250)
                y += 1;
                if (p.step_reset) ld = 0;
            }
            pc.y = y;
            if (lane == 2) sw += 1;  // time (ref:253)
            if ((cm >> (y + 1)) & 1) {                           // _has_dropped (ref:202-203)
                ld += 1;
                if (ld >= p.lock_mod) ld %= p.lock_mod;          // ref:258
                if (ld == 0) {
                    cols_lock<ColT>(col, sw, pc, cl, lane, X3, on_board, p, e, reward, done, errbits);
                    if (!done) cl = cols_cells(pc.id, 0);        // the fresh piece
                }
            }
            put(sw, lane, 1, ld);
            pm = cols_piece<ColT>(cl, pc.x, pc.y, X3, H);
        }
        put(sw, lane, 0, pack_piece(pc));
        if (has_info && lane < kStateWords) *info_p = lane == 0 ? pc.id : sw;  // get_info (ref:232-241), pre-reset
        ColT shown = on_board ? (col | pm) : (ColT)0;                         // _set_piece(True), copy (ref:301-302)
        if (on_board) col &= ~pm;                                             // _set_piece(False) (ref:303), literally
        bool has_term = false;
        ColT term = 0;
        if (done) {
            if (p.stats && lane >= 2 && lane <= 4)  // sum(time), sum(score), sum(lines) at done
                atomicAdd(p.stats + (lane == 2 ? 1 : lane == 3 ? 3 : 2), (unsigned long long)(long long)sw);
            if (p.stats && lane == 0) atomicAdd(p.stats, 1ull);
            if (p.auto_reset) {  // VecEnv: the returned observation is the reset one — empty board, piece not drawn
                if (want_term) { has_term = true; term = shown; }
                if (lane >= 2 && lane <= 6) sw = 0;  // clear() (ref:306-315)
                pc.id = cols_spawn(sw, lane, p, e, errbits);
                pc.rot = 0; pc.x = W / 2; pc.y = 0;
                put(sw, lane, 0, pack_piece(pc));
                if (on_board) col = 0;
                shown = 0;
            }
        }
        if (lane == 0) {
            *rew_p = (float)reward;
            *done_p = (unsigned char)done;
        }
        if (fast_obs) {
            const uint32_t c0 = (uint32_t)(warp_get<ColT>(shown, src0) >> sh0);
            if (lane < nq) reinterpret_cast<float4 *>(obs_p)[lane] = s_lut[c0 & 15u];
            if (nq > 32) {
                const uint32_t c1 = (uint32_t)(warp_get<ColT>(shown, src1) >> sh1);
                if (lane + 32 < nq) reinterpret_cast<float4 *>(obs_p)[lane + 32] = s_lut[c1 & 15u];
            }
        } else if (has_obs) {
            cols_write_ram<ColT>(shown, obs_p, s_lut, p, lane);
        }
        if (has_term) cols_write_ram<ColT>(term, term_p, s_lut, p, lane);
        if (t + 1 < p.T) {  // next step of st_step_many
            act_p += n; rew_p += n; done_p += n;
            info_p += p.info_t_stride;
            obs_p += obs_step;
            term_p += obs_step;
            action = (int)__reduce_or_sync(FULL, next_u);
        }
    }
    if (lane < kStateWords) rec_w[lane] = (uint32_t)sw;
    if (on_board) {
        if constexpr (CW == 1) {
            rec_w[kStateWords + X] = (uint32_t)col;
        } else {
            rec_w[kStateWords + 2 * X] = (uint32_t)col;
            rec_w[kStateWords + 2 * X + 1] = (uint32_t)((unsigned long long)col >> 32);
        }
    }
    if (errbits && p.err && lane == 0) atomicOr(p.err, errbits);
}

// ram observations, step launches (single or T steps), boards up to 24 columns: everything else stays on K1
static bool cols_eligible(const Params &p, int obs_type)
{
    return obs_type == 0 && p.mode == MODE_STEP && p.n > 0 && p.W <= kColsMaxW && p.H <= 63;
}

// Envs one wave of an instantiation holds on this device (resident CTAs per SM x SMs x envs per CTA), cached.
template <typename ColT, int MINB>
static long long cols_wave_envs()
{
    static long long cap[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 0;
    if (!cap[dev]) {
        int ctas = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, st_step_cols_kernel<ColT, MINB>, 32 * kColsWarpsPerCta, 0) != cudaSuccess)
            ctas = 0;
        cap[dev] = (long long)ctas * tpe_sm_count() * kColsWarpsPerCta;
        if (!cap[dev]) cap[dev] = -1;
    }
    return cap[dev] > 0 ? cap[dev] : 0;
}

template <typename ColT>
static cudaError_t launch_cols_t(const Params &p, cudaLaunchConfig_t &cfg)
{
    // register budgets, richest first: 1 (no cap: ~72 registers), 7 (<= 73), 0 (the compiler's own choice: ~53)
    const int force = env_int("ST_B200_COLS_MINB", -1);  // experiment knob, read at every launch like ST_B200_RAM_PATH
    const int pick = force >= 0 ? force : p.n <= cols_wave_envs<ColT, 1>() ? 1 : p.n <= cols_wave_envs<ColT, 7>() ? 7 : 0;
    if (pick == 1) return cudaLaunchKernelEx(&cfg, st_step_cols_kernel<ColT, 1>, p);
    if (pick == 7) return cudaLaunchKernelEx(&cfg, st_step_cols_kernel<ColT, 7>, p);
    return cudaLaunchKernelEx(&cfg, st_step_cols_kernel<ColT, 0>, p);
}

static cudaError_t launch_cols(const Params &p, cudaStream_t stream)
{
    static const bool pdl = getenv("ST_B200_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((p.n + kColsWarpsPerCta - 1) / kColsWarpsPerCta));
    cfg.blockDim = dim3(32 * kColsWarpsPerCta);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    count_launch();
    if (p.col_words == 1) return launch_cols_t<uint32_t>(p, cfg);
    return launch_cols_t<unsigned long long>(p, cfg);
}

}  // namespace st
