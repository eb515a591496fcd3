// K1b — thread-per-env ram step on COLUMN bitboards (included by st_kernels.cu).
//
// The warp-per-env kernel (K1) spends ~400 warp-instructions per env-step: ideal for small batches, where the
// machine is latency-bound and a whole warp per env keeps every rare branch uniform, but issue-bound from
// ~16k envs up.  Here one THREAD steps one env and the warp does the memory work cooperatively:
//   1. the warp copies its env records (contiguous in HBM) into shared memory with 16-byte loads; the record
//      pitch in smem is odd, so lane r touching word c of ITS record is conflict-free;
//   2. each lane runs TetrisEngine.step (ref:243-304) on its record in smem.  The record holds the board as one
//      word per COLUMN (bit y of column x = cell (x, y)), which is what makes the per-thread engine short:
//        * a piece is four cells (i, j); the collision answer for EVERY anchor height at once is the OR over the
//          four cells of column[x + i] shifted by j (ref:29-36: cells above the board fall off the low end of the
//          shift and are exempt from board and walls alike; a column outside the board is all ones), plus the floor;
//          soft drop, hard drop (one ctz — no search loop), gravity and the grounded test read bits of that mask;
//        * full rows are the AND of all columns, holes are H - top - popc per column, height is popc of the OR;
//          a cleared row is squeezed out of every column with three logic ops;
//   3. info / reward / done go out coalesced, auto-reset envs are cleared, every other lane ORs its piece
//      into its smem columns (the reference's _set_piece(True), ref:301);
//   4. the warp expands all its boards into float32 [W][H] — the observation is column-major like the record, four
//      cells are four adjacent bits of one word — with 16-byte stores, 512 contiguous bytes per warp instruction;
//   5. lanes erase their piece again (ref:303) and the records are stored back, coalesced.
#pragma once

namespace st {

#ifndef ST_TPE_WARPS
#define ST_TPE_WARPS 4
#endif
constexpr int kTpeWarps = ST_TPE_WARPS;

// Piece cells for the thread-per-env engine: entry = four cells, byte c = (i + 3) | (j + 3) << 3 (ref:10-19, rotated
// as ref:22-26), then maxj + 3 in bits 32..35.
struct CellTab {
    unsigned long long e[28];
};

constexpr CellTab make_cell_tab()
{
    const int base[7][4][2] = {
        {{0, 0}, {-1, 0}, {1, 0}, {0, -1}},    // T
        {{0, 0}, {-1, 0}, {0, -1}, {0, -2}},   // J
        {{0, 0}, {1, 0}, {0, -1}, {0, -2}},    // L
        {{0, 0}, {-1, 0}, {0, -1}, {1, -1}},   // Z
        {{0, 0}, {-1, -1}, {0, -1}, {1, 0}},   // S
        {{0, 0}, {0, -1}, {0, -2}, {0, -3}},   // I
        {{0, 0}, {0, -1}, {-1, 0}, {-1, -1}},  // O
    };
    CellTab t{};
    for (int id = 0; id < 7; ++id) {
        int c[4][2] = {};
        for (int k = 0; k < 4; ++k) { c[k][0] = base[id][k][0]; c[k][1] = base[id][k][1]; }
        for (int r = 0; r < 4; ++r) {
            int mx = -3;
            unsigned long long m = 0;
            for (int k = 0; k < 4; ++k) {
                mx = c[k][1] > mx ? c[k][1] : mx;
                m |= (unsigned long long)((c[k][0] + 3) | ((c[k][1] + 3) << 3)) << (8 * k);
            }
            m |= (unsigned long long)(mx + 3) << 32;
            t.e[id * 4 + r] = m;
            for (int k = 0; k < 4; ++k) { int i = c[k][0], j = c[k][1]; c[k][0] = j; c[k][1] = -i; }
        }
    }
    return t;
}

__constant__ CellTab c_cells = make_cell_tab();

template <typename ColT> struct ColOps;
template <> struct ColOps<uint32_t> {
    static constexpr int kWords = 1;
    static __device__ __forceinline__ int popc(uint32_t v) { return __popc(v); }
    static __device__ __forceinline__ int ffs(uint32_t v) { return __ffs((int)v); }
};
template <> struct ColOps<unsigned long long> {
    static constexpr int kWords = 2;
    static __device__ __forceinline__ int popc(unsigned long long v) { return __popcll(v); }
    static __device__ __forceinline__ int ffs(unsigned long long v) { return __ffsll((long long)v); }
};

// One env record in shared memory: 15 scalar words, then W column words (two 32-bit halves when H > 31).
template <typename ColT>
struct TpeRec {
    uint32_t *w;
    __device__ __forceinline__ ColT col(int X) const
    {
        if constexpr (ColOps<ColT>::kWords == 1) return (ColT)w[kStateWords + X];
        else return (ColT)w[kStateWords + 2 * X] | ((ColT)w[kStateWords + 2 * X + 1] << 32);
    }
    __device__ __forceinline__ void set_col(int X, ColT v) const
    {
        if constexpr (ColOps<ColT>::kWords == 1) {
            w[kStateWords + X] = (uint32_t)v;
        } else {
            w[kStateWords + 2 * X] = (uint32_t)v;
            w[kStateWords + 2 * X + 1] = (uint32_t)((unsigned long long)v >> 32);
        }
    }
};

struct Cells {
    int i[4], j[4];
    int maxj;
};

__device__ __forceinline__ Cells tpe_cells(const unsigned long long *s_cells, int id, int rot)
{
    const unsigned long long e = s_cells[id * 4 + rot];  // smem copy: lanes index different entries
    const uint32_t lo = (uint32_t)e;
    Cells c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c.i[k] = (int)((lo >> (8 * k)) & 7u) - 3;
        c.j[k] = (int)((lo >> (8 * k + 3)) & 7u) - 3;
    }
    c.maxj = (int)((uint32_t)(e >> 32) & 15u) - 3;
    return c;
}

// Bit y' set <=> is_occupied(shape, (x, y'), board) (ref:29-36), for every anchor height y' at once.
template <typename ColT>
__device__ __forceinline__ ColT tpe_collisions(const TpeRec<ColT> &rec, const Cells &c, int x, int W, int H)
{
    ColT cm = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int X = x + c.i[k];
        const ColT v = (unsigned)X < (unsigned)W ? rec.col(X) : ~(ColT)0;  // outside the board: ref:34
        // cell row = y' + j: the anchors it blocks are the column shifted by j; for j < 0 the low -j anchors put the
        // cell above the board, where nothing is tested (ref:32-33) — the left shift leaves exactly those bits clear
        cm |= c.j[k] >= 0 ? (v >> c.j[k]) : (v << -c.j[k]);
    }
    int fl = H - c.maxj;  // anchors whose lowest cell is at or below the floor (ref:34 `y >= board.shape[1]`)
    fl = fl < 0 ? 0 : fl;
    cm |= ~(ColT)0 << fl;
    return cm;
}

// _new_piece / _choose_shape (ref:183-200) on the record's shape_counts (words 8..14).
template <typename ColT>
__device__ __forceinline__ int tpe_spawn(const TpeRec<ColT> &rec, const Params &p, int e, int &errbits)
{
    int c[7];
    int total = 0, mx = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        c[i] = (int)rec.w[8 + i];
        total += c[i];
        mx = c[i] > mx ? c[i] : mx;
    }
    int id;
    if (p.queue) {
        int k = total;
        if (k >= p.queue_len) { errbits |= 1; k %= p.queue_len; }
        id = p.queue[(size_t)e * (unsigned)p.queue_len + k] % 7;
    } else {
        const int S = 35 + 7 * mx - total;
        const uint32_t u = philox_draw(p.seed_lo, p.seed_hi, (unsigned long long)(p.env_id_base + e), (uint32_t)total);
        const int r = 1 + (int)__umulhi(u, (uint32_t)S);
        int acc = 0;
        id = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            acc += 5 + mx - c[i];
            id += (r > acc) ? 1 : 0;
        }
    }
    rec.w[8 + id] += 1u;
    return id;
}

// The lock branch (ref:262-299).
template <typename ColT>
__device__ __forceinline__ void tpe_lock(const TpeRec<ColT> &rec, Piece &pc, const Cells &cl, const Params &p, int e,
                                         int &reward, int &done, int &errbits)
{
    using Ops = ColOps<ColT>;
    const int H = p.H, W = p.W;
    const ColT hmask = ~(ColT)0 >> (8 * (int)sizeof(ColT) - H);
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // _set_piece(True) (ref:263): in-board cells only
        const int X = pc.x + cl.i[k], Y = pc.y + cl.j[k];
        if (Y >= 0 && Y < H && (unsigned)X < (unsigned)W) rec.set_col(X, rec.col(X) | ((ColT)1 << Y));
    }
    // One pass over the columns answers _clear_lines' can_clear (ref:206: a row is full when every column has it),
    // _count_holes (ref:218-220: per column, the empty cells below its top-most filled one = H - top - popc) and
    // sum(np.any(board, axis=0)) (ref:287,289: rows with any cell = popc of the OR of all columns).
    ColT full = hmask, any = 0;
    int holes = 0;
    for (int x = 0; x < W; ++x) {
        const ColT c = rec.col(x);
        full &= c;
        any |= c;
        holes += c ? H + 1 - Ops::ffs(c) - Ops::popc(c) : 0;
    }
    const int k = Ops::popc(full);
    if (k) {  // rare: squeeze the full rows out of every column, top-most first (ref:207-214), then recount
        any = 0;
        holes = 0;
        for (int x = 0; x < W; ++x) {
            ColT c = rec.col(x);
            ColT f = full;
            while (f) {
                const ColT bit = f & (~f + 1);  // lowest set bit = top-most full row r
                f ^= bit;
                c = (c & ~(bit | (bit - 1))) | ((c & (bit - 1)) << 1);  // rows below r stay, rows above move down by one
            }
            rec.set_col(x, c);
            any |= c;
            holes += c ? H + 1 - Ops::ffs(c) - Ops::popc(c) : 0;
        }
        rec.w[4] += (uint32_t)k;
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
    rec.w[3] += (uint32_t)dscore;
    const int old_holes = (int)rec.w[5];
    rec.w[5] = (uint32_t)holes;
    if (any & 1) {  // np.any(board[:, 0]) (ref:277-281)
        rec.w[7] += 1u;
        done = 1;
        reward = -100;
    } else {
        if (p.pen_height) {
            reward -= nonempty;
        } else if (p.pen_height_inc) {
            const int ph = (int)rec.w[6];
            if (nonempty > ph) reward -= 10 * (nonempty - ph);
            rec.w[6] = (uint32_t)nonempty;
        }
        if (p.pen_holes) reward -= 5 * holes;
        else if (p.pen_holes_inc) reward -= 5 * (holes - old_holes);
        pc.id = tpe_spawn(rec, p, e, errbits);  // ref:299
        pc.rot = 0; pc.x = W / 2; pc.y = 0;
    }
}

// TetrisEngine.step (ref:243-304) up to, not including, the composition of the returned state.
template <typename ColT>
__device__ __forceinline__ void tpe_engine_step(const TpeRec<ColT> &rec, int action, const Params &p, int e,
                                                const unsigned long long *s_cells, int &reward, int &done, int &errbits)
{
    const int H = p.H, W = p.W;
    Piece pc = unpack_piece((int)rec.w[0]);
    reward = p.reward_step;
    done = 0;
    if (pc.id >= 7) { errbits |= 4; return; }
    if (action > 6) { errbits |= 2; action = 6; }
    int r2 = pc.rot, x2 = pc.x;
    if (action == 0) x2 -= 1;
    if (action == 1) x2 += 1;
    if (action == 4) r2 = (r2 + 1) & 3;
    if (action == 5) r2 = (r2 + 3) & 3;
    Cells cl = tpe_cells(s_cells, pc.id, r2);
    ColT cm = tpe_collisions(rec, cl, x2, W, H);
    const bool moved = (r2 != pc.rot) || (x2 != pc.x);
    if (moved && ((cm >> pc.y) & 1)) {  // blocked: stay (ref:41,46,64,69)
        cl = tpe_cells(s_cells, pc.id, pc.rot);
        cm = tpe_collisions(rec, cl, pc.x, W, H);
    } else {
        pc.rot = r2; pc.x = x2;
    }
    int y = pc.y;
    if (action == 3 && !((cm >> (y + 1)) & 1)) y += 1;  // soft_drop (ref:49-51)
    if (action == 2) {                                   // hard_drop (ref:54-59): first blocked height below
        const ColT below = cm >> (y + 1);
        if (below) y += ColOps<ColT>::ffs(below) - 1;
    }
    int ld = (int)rec.w[1];
    if (!((cm >> (y + 1)) & 1)) {                        // gravity (ref:247-250)
        y += 1;
        if (p.step_reset) ld = 0;
    }
    pc.y = y;
    rec.w[2] += 1u;  // time (ref:253)
    if ((cm >> (y + 1)) & 1) {                           // _has_dropped (ref:202-203)
        ld += 1;
        if (ld >= p.lock_mod) ld %= p.lock_mod;
        if (ld == 0) tpe_lock(rec, pc, cl, p, e, reward, done, errbits);
    }
    rec.w[1] = (uint32_t)ld;
    rec.w[0] = (uint32_t)pack_piece(pc);
}

template <typename ColT>
__global__ void __launch_bounds__(32 * kTpeWarps) st_step_tpe_kernel(const __grid_constant__ Params p)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];
    __shared__ unsigned long long s_cells[28];
    constexpr int CW = ColOps<ColT>::kWords;
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 28) s_cells[threadIdx.x] = c_cells.e[threadIdx.x];
    const int H = p.H, W = p.W;
    const int SW = p.stride >> 2;  // words per record in HBM
    const int pitch = SW | 1;      // odd pitch in smem: lane r, word c -> bank (r * pitch + c) % 32, conflict-free
    const int epw = p.tpe_epw;     // envs per warp (32, 16, 8 or 4): fewer envs per warp = more warps for mid-size batches
    uint32_t *recs = s_dyn + warp * epw * pitch;
    const long long e0 = ((long long)blockIdx.x * kTpeWarps + warp) * epw;
    int nvalid = (int)(p.n - e0 < epw ? p.n - e0 : epw);
    nvalid = nvalid < 0 ? 0 : nvalid;
    __syncthreads();  // s_cells
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (nvalid == 0) return;

    // the action bytes are issued first so that their miss overlaps the record copy
    const int e = (int)e0 + lane;
    unsigned int action_u = 6u;
    if (lane < nvalid) asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(action_u) : "l"(p.actions + e));

    // 1. records HBM -> smem
    uint32_t *g_rec = reinterpret_cast<uint32_t *>(p.state + e0 * (long long)p.stride);
    const int nwords = nvalid * SW;
    const bool vec_ok = pitch == SW && ((e0 * (long long)p.stride) & 15) == 0;  // contiguous in both, 16-byte aligned
    if (vec_ok) {
        const int nvec = nwords >> 2;
        for (int i = lane; i < nvec; i += 32) reinterpret_cast<uint4 *>(recs)[i] = reinterpret_cast<const uint4 *>(g_rec)[i];
        for (int i = (nvec << 2) + lane; i < nwords; i += 32) recs[i] = g_rec[i];
    } else {
        for (int i = lane; i < nwords; i += 32) {
            const int r = (int)(((uint32_t)i * p.inv_sw20) >> 20);
            recs[r * pitch + (i - r * SW)] = g_rec[i];
        }
    }
    __syncwarp();

    const TpeRec<ColT> rec = {recs + lane * pitch};
    int errbits = 0;
    const size_t n_envs = (size_t)p.n;
    for (int t = 0; t < p.T; ++t) {  // st_step_many: the records stay in shared memory between steps
    // 2. engine, one env per lane
    int reward = 0, done = 0;
    if (lane < nvalid) tpe_engine_step(rec, (int)action_u, p, e, s_cells, reward, done, errbits);
    if (t + 1 < p.T && lane < nvalid)  // next step's action, in flight during the cooperative phases
        asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(action_u) : "l"(p.actions + (size_t)(t + 1) * n_envs + e));
    __syncwarp();

    // 3. info (pre-reset), reward, done; then auto-reset or piece overlay
    if (p.info) {
        int32_t *g_info = p.info + (long long)t * p.info_t_stride + e0 * kStateWords;
        for (int j = lane; j < nvalid * kStateWords; j += 32) {
            const int r = (j * 4370) >> 16;  // j / 15 for j < 480
            const int c = j - r * kStateWords;
            const uint32_t v = recs[r * pitch + c];
            g_info[j] = c == 0 ? (int32_t)(v & 15u) : (int32_t)v;
        }
    }
    if (lane < nvalid) {
        p.reward[(size_t)t * n_envs + e] = (float)reward;
        p.done[(size_t)t * n_envs + e] = (unsigned char)done;
        if (done && p.stats) {
            atomicAdd(p.stats, 1ull);
            atomicAdd(p.stats + 1, (unsigned long long)(long long)(int)rec.w[2]);
            atomicAdd(p.stats + 2, (unsigned long long)(long long)(int)rec.w[4]);
            atomicAdd(p.stats + 3, (unsigned long long)(long long)(int)rec.w[3]);
        }
    }
    __syncwarp();
    const bool u8 = p.obs_u8 != 0;
    const int nel = W * H;
    if (p.term_obs) {  // terminal observation of the envs that end here: their board already holds the locked piece
        unsigned term = __ballot_sync(FULL, lane < nvalid && done && p.auto_reset);
        char *tb = reinterpret_cast<char *>(p.term_obs) + ((long long)t * p.obs_t_stride + e0 * (long long)p.obs_elems) * (u8 ? 1 : 4);
        while (term) {
            const int r = __ffs((int)term) - 1;
            term &= term - 1;
            const uint32_t *cw = recs + r * pitch + kStateWords;
            char *dst = tb + (size_t)r * nel * (u8 ? 1 : 4);
            for (int i = lane; i < nel; i += 32) {
                const int x = (int)(((uint32_t)i * p.inv_h20) >> 20);
                const int yy = i - x * H;
                const bool on = ((cw[CW * x + (yy >> 5)] >> (yy & 31)) & 1u) != 0u;
                if (u8) reinterpret_cast<unsigned char *>(dst)[i] = on ? 1 : 0;
                else reinterpret_cast<float *>(dst)[i] = on ? 1.0f : 0.0f;
            }
        }
        __syncwarp();
    }
    int pX[4] = {-1, -1, -1, -1}, pY[4] = {0, 0, 0, 0};  // the piece's in-board cells (column, row); -1 = none
    if (lane < nvalid) {
        if (done && p.auto_reset) {  // clear() (ref:306-315): the reset observation is the empty board
#pragma unroll
            for (int i = 2; i <= 6; ++i) rec.w[i] = 0u;
            Piece pc;
            pc.id = tpe_spawn(rec, p, e, errbits);
            pc.rot = 0; pc.x = W / 2; pc.y = 0;
            rec.w[0] = (uint32_t)pack_piece(pc);
            for (int i = 0; i < W * CW; ++i) rec.w[kStateWords + i] = 0u;
        } else {
            const Piece pc = unpack_piece((int)rec.w[0]);
            if (pc.id < 7) {  // _set_piece(True) (ref:301)
                const Cells cl = tpe_cells(s_cells, pc.id, pc.rot);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int X = pc.x + cl.i[k], Y = pc.y + cl.j[k];
                    if (Y >= 0 && Y < H && (unsigned)X < (unsigned)W) {
                        pX[k] = X; pY[k] = Y;
                        rec.set_col(X, rec.col(X) | ((ColT)1 << Y));
                    }
                }
            }
        }
    }
    __syncwarp();

    // 4. observations: float32 [W][H] per env (ref:421-424, 400); 16-byte stores when H % 4 == 0, else 4-byte ones
    if (p.obs && (H & 3) != 0) {
        char *o = reinterpret_cast<char *>(p.obs) + ((long long)t * p.obs_t_stride + e0 * (long long)p.obs_elems) * (u8 ? 1 : 4);
        for (int i0 = 0; i0 < nel; i0 += 32) {
            const int i = i0 + lane;
            const int x = (int)(((uint32_t)i * p.inv_h20) >> 20);
            const int yy = i - x * H;
            if (i < nel) {
                const int wi = kStateWords + CW * x + (yy >> 5), sh = yy & 31;
                for (int r = 0; r < nvalid; ++r) {
                    const bool on = ((recs[r * pitch + wi] >> sh) & 1u) != 0u;
                    if (u8) reinterpret_cast<unsigned char *>(o)[r * nel + i] = on ? 1 : 0;
                    else reinterpret_cast<float *>(o)[r * nel + i] = on ? 1.0f : 0.0f;
                }
            }
        }
    } else if (p.obs) {
        const int hq = H >> 2, nq = W * hq;
        char *o4 = reinterpret_cast<char *>(p.obs) + ((long long)t * p.obs_t_stride + e0 * (long long)p.obs_elems) * (u8 ? 1 : 4);
        for (int q0 = 0; q0 < nq; q0 += 32) {
            const int q = q0 + lane;
            const int x = (int)(((uint32_t)q * p.inv_hq20) >> 20);
            const int yq = q - x * hq;
            if (q < nq) {  // cells (x, 4 yq .. 4 yq + 3) are four adjacent bits of column x
                const int wi = kStateWords + CW * x + (yq >> 3), sh = 4 * (yq & 7);
                for (int r = 0; r < nvalid; ++r) {
                    const uint32_t b = recs[r * pitch + wi] >> sh;
                    const float4 v = make_float4((b & 1u) ? 1.0f : 0.0f, (b & 2u) ? 1.0f : 0.0f, (b & 4u) ? 1.0f : 0.0f,
                                                 (b & 8u) ? 1.0f : 0.0f);
                    store4(o4, (size_t)(r * nq + q), v, u8);
                }
            }
        }
    }
    __syncwarp();

    // 5. _set_piece(False) (ref:303), literally: the cells of the piece are cleared on the board
    if (lane < nvalid) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (pX[k] >= 0) rec.set_col(pX[k], rec.col(pX[k]) & ~((ColT)1 << pY[k]));
    }
    __syncwarp();
    }  // for t
    if (vec_ok) {
        const int nvec = nwords >> 2;
        for (int i = lane; i < nvec; i += 32) reinterpret_cast<uint4 *>(g_rec)[i] = reinterpret_cast<const uint4 *>(recs)[i];
        for (int i = (nvec << 2) + lane; i < nwords; i += 32) g_rec[i] = recs[i];
    } else {
        for (int i = lane; i < nwords; i += 32) {
            const int r = (int)(((uint32_t)i * p.inv_sw20) >> 20);
            g_rec[i] = recs[r * pitch + (i - r * SW)];
        }
    }
    errbits = __reduce_or_sync(FULL, errbits);
    if (errbits && p.err && lane == 0) atomicOr(p.err, errbits);
}

template <typename ColT>
static cudaError_t launch_tpe_t(const Params &p, cudaStream_t stream)
{
    const long long nwarps = (p.n + p.tpe_epw - 1) / p.tpe_epw;
    const long long nctas = (nwarps + kTpeWarps - 1) / kTpeWarps;
    const int pitch = (p.stride >> 2) | 1;
    const size_t smem = (size_t)kTpeWarps * p.tpe_epw * pitch * 4;
    static const bool pdl = getenv("ST_B200_NO_PDL") == nullptr;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(st_step_tpe_kernel<ColT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)nctas);
    cfg.blockDim = dim3(32 * kTpeWarps);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    count_launch();
    return cudaLaunchKernelEx(&cfg, st_step_tpe_kernel<ColT>, p);
}

// Round-1 measurements of the row-bitboard version of this kernel (tools/ram_path_sweep.py, us per step; 10x20 /
// 20 wide x 40 high boards) — kept as the baseline the column version is compared with in DESIGN.md:
//   n        warp    8/warp  16/warp  32/warp   |   warp    8/warp  16/warp  32/warp
//   8192     8.2     12.5    11.9     11.2      |   10.7    22.3    20.9     21.3
//   65536    34.8    22.6    20.5     27.7      |   50.8    49.1    59.2     73.4
//   1048576  487.5   253.6   186.4    246.5     |  733.1   652.9   703.2    738.4
static long long tpe_min_envs(const Params &p) { return p.H <= 31 ? 24576 : 65536; }
static int tpe_default_epw(const Params &p) { return (p.H <= 31 && p.n >= 49152) ? 16 : 8; }

// Thread-per-env path: single-step ram launches.
static bool tpe_eligible(const Params &p, int obs_type)
{
    return obs_type == 0 && p.mode == MODE_STEP && p.n > 0 && (p.obs_t_stride & 3) == 0;
}

static cudaError_t launch_tpe(const Params &p0, cudaStream_t stream)
{
    Params p = p0;
    // envs per warp: the per-warp engine chain is the same for 8 or 32 envs, so mid-size batches get more warps
    const char *ov = getenv("ST_B200_TPE_EPW");
    p.tpe_epw = ov ? atoi(ov) : tpe_default_epw(p);
    if (p.tpe_epw != 4 && p.tpe_epw != 8 && p.tpe_epw != 16 && p.tpe_epw != 32) p.tpe_epw = 32;
    if (p.col_words == 1) return launch_tpe_t<uint32_t>(p, stream);
    return launch_tpe_t<unsigned long long>(p, stream);
}

}  // namespace st
