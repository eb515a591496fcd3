// K1b — thread-per-env ram step on COLUMN bitboards (included by st_kernels.cu).
//
// A warp per env costs 270-580 warp-instructions per env-step: ideal for small batches, where the machine is
// latency-bound and a whole warp per env keeps every rare branch uniform, but issue-bound from ~14k envs up.  Here one
// THREAD steps one env and the warp does the memory work cooperatively, for one GROUP of `epw` consecutive envs at a
// time (a warp walks over several groups when the grid is capped):
//   1. the group's env records (contiguous in HBM) arrive in shared memory through cp.async (LDGSTS); the record pitch
//      in smem is odd, so lane r touching word c of ITS record is conflict-free;
//   2. each lane runs TetrisEngine.step (ref:243-304) on its record in smem.  The record holds the board as one
//      word per COLUMN (bit y of column x = cell (x, y)), which is what makes the per-thread engine short:
//        * a piece is four cells (i, j); the collision answer for EVERY anchor height at once is the OR over the
//          four cells of column[x + i] shifted by j (ref:29-36: cells above the board fall off the low end of the
//          shift and are exempt from board and walls alike; a column outside the board is all ones), plus the floor;
//          soft drop, hard drop (one ctz — no search loop), gravity and the grounded test read bits of that mask;
//        * full rows are the AND of all columns, holes are H - top - popc per column, height is popc of the OR;
//          a cleared row is squeezed out of every column with three logic ops;
//   3. reward / done go out, auto-reset envs are cleared, every other lane ORs its piece into its smem columns (the
//      reference's _set_piece(True), ref:301);
//   4. the warp expands the group's boards into float32 [W][H] — the observation is column-major like the record —
//      with direct 16-byte stores, 512 contiguous bytes per warp instruction (tpe_obs_direct); boards whose height is
//      not a multiple of 4 go through two chunk buffers in shared memory and cp.async.bulk (TMA) stores instead;
//   5. info, then lanes erase their piece again (ref:303) and the records are stored back, coalesced.
// DESIGN.md section 3 (K1b) has the measured phase timeline and what was tried against it.
#pragma once

namespace st {

#ifndef ST_TPE_MINBLOCKS
#define ST_TPE_MINBLOCKS 1  // 1024-thread CTAs x 1 = 64 registers per thread: 32 resident warps per SM
#endif
constexpr int kTpeMaxWarps = 32;  // warps per CTA is a launch-time choice (1..32)

// Phase timestamps of every warp (profiling builds only: -DST_TPE_TRACE=1, tools/tpe_trace.py)
#ifndef ST_TPE_TRACE
#define ST_TPE_TRACE 0
#endif
#if ST_TPE_TRACE
__device__ unsigned long long *g_tpe_trace = nullptr;  // [8 launches][kTraceWarps][16]
constexpr size_t kTraceWarps = 1 << 15;
__device__ __forceinline__ unsigned long long tpe_now()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define TPE_MARK(slot)                                                                                        \
    do {                                                                                                      \
        if (g_tpe_trace && lane == 0 && g == (int)blockIdx.x * wpc + warp && (size_t)g < kTraceWarps)           \
            g_tpe_trace[((size_t)(p.draw_piece & 7) * kTraceWarps + (size_t)blockIdx.x * wpc + warp) * 16 + (slot)] = tpe_now(); \
    } while (0)
#else
#define TPE_MARK(slot) do { } while (0)
#endif

// Piece tables for the thread-per-env engine (ref:10-19, rotated as ref:22-26).
//   e[s]: the four cells, byte c = (i + 3) | (j + 3) << 3, then maxj + 3 in bits 32..35          (collision test)
//   c[s]: the same piece by COLUMN, four 10-bit fields = (i + 3) | pattern << 3, bit j + 3 of the pattern = cell
//         (i, j); unused fields are 0                                                              (drawing the piece)
struct CellTab {
    unsigned long long e[28];
    unsigned long long c[28];
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
            unsigned long long cols = 0;
            int nf = 0;
            for (int i = -3; i <= 3; ++i) {
                unsigned pat = 0;
                for (int k = 0; k < 4; ++k)
                    if (c[k][0] == i) pat |= 1u << (c[k][1] + 3);
                if (pat) cols |= (unsigned long long)((unsigned)(i + 3) | (pat << 3)) << (10 * nf++);
            }
            t.c[id * 4 + r] = cols;
            for (int k = 0; k < 4; ++k) { int i = c[k][0], j = c[k][1]; c[k][0] = j; c[k][1] = -i; }
        }
    }
    return t;
}

// read once per CTA into shared memory, one entry per thread: plain global memory (a lane-indexed __constant__
// load would be replayed once per distinct address)
__device__ const CellTab g_cells = make_cell_tab();

template <typename ColT> struct ColOps;
template <> struct ColOps<uint32_t> {
    static constexpr int kWords = 1;
    static __device__ __forceinline__ int popc(uint32_t v) { return __popc(v); }
    static __device__ __forceinline__ int ffs(uint32_t v) { return __ffs((int)v); }
    // column v moved by j = s3 - 3 rows: bit y' of the result = bit y' + j of v (bits shifted in are clear)
    static __device__ __forceinline__ uint32_t by_rows(uint32_t v, int s3) { return (uint32_t)(((unsigned long long)v << 3) >> s3); }
};
template <> struct ColOps<unsigned long long> {
    static constexpr int kWords = 2;
    static __device__ __forceinline__ int popc(unsigned long long v) { return __popcll(v); }
    static __device__ __forceinline__ int ffs(unsigned long long v) { return __ffsll((long long)v); }
    static __device__ __forceinline__ unsigned long long by_rows(unsigned long long v, int s3)
    {
        return s3 >= 3 ? (v >> (s3 - 3)) : (v << (3 - s3));
    }
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
    int i[4];   // column offset of cell k
    int s3[4];  // row offset + 3
    int maxj;
};

// The piece at anchor (x, y), column by column: up to four board columns X[c] (-1 = none / outside the board) and
// the piece's cells in that column as a mask of board rows (rows above the board and below the floor drop out, which
// is _set_piece's `0 <= y < height` test, ref:325-326).
template <typename ColT>
struct PieceCols {
    int X[4];
    ColT m[4];
};

template <typename ColT>
__device__ __forceinline__ PieceCols<ColT> tpe_piece_cols(const unsigned long long *s_cols, int id, int rot, int x, int y, int W, int H)
{
    const unsigned long long e = s_cols[id * 4 + rot];
    const ColT hmask = (((ColT)1 << H) - 1);
    PieceCols<ColT> pc;
    constexpr int kYMax = sizeof(ColT) == 4 ? 56 : 66;  // an injected anchor far below the floor: every cell drops
    y = y > kYMax ? kYMax : y;                          // out, no shift overflows
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const uint32_t f = (uint32_t)(e >> (10 * c)) & 1023u;
        const int X = x + (int)(f & 7u) - 3;
        const uint32_t pat = f >> 3;
        ColT m;
        if constexpr (sizeof(ColT) == 4) m = (ColT)(((unsigned long long)pat << y) >> 3);
        else m = y >= 3 ? ((ColT)pat << (y - 3)) : ((ColT)pat >> (3 - y));  // shift <= 63
        pc.m[c] = m & hmask;
        pc.X[c] = (pat != 0u && (unsigned)X < (unsigned)W) ? X : -1;
    }
    return pc;
}

__device__ __forceinline__ Cells tpe_cells(const unsigned long long *s_cells, int id, int rot)
{
    const unsigned long long e = s_cells[id * 4 + rot];  // smem copy: lanes index different entries
    const uint32_t lo = (uint32_t)e;
    Cells c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        c.i[k] = (int)((lo >> (8 * k)) & 7u) - 3;
        c.s3[k] = (int)((lo >> (8 * k + 3)) & 7u);
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
        // cell row = y' + j: the anchors it blocks are the column moved by j rows; for j < 0 the low -j anchors put
        // the cell above the board, where nothing is tested (ref:32-33) — the shift leaves exactly those bits clear
        cm |= ColOps<ColT>::by_rows(v, c.s3[k]);
    }
    int fl = H - c.maxj;  // anchors whose lowest cell is at or below the floor (ref:34 `y >= board.shape[1]`)
    fl = fl < 0 ? 0 : fl;
    cm |= ~(ColT)0 << fl;
    return cm;
}

// _new_piece / _choose_shape (ref:183-200) on the record's shape_counts (words 8..14).
template <typename ColT>
__device__ __forceinline__ int tpe_spawn(const TpeRec<ColT> &rec, const Params &p, long long e, int &errbits)
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

// Over all columns: the rows every column has (full), the rows any column has, and _count_holes (ref:218-220: per
// column, the empty cells below its top-most filled one = H - top - popc; an empty column finds the sentinel bit H).
template <typename ColT>
__device__ __forceinline__ void tpe_scan(const TpeRec<ColT> &rec, int W, int H, ColT &full, ColT &any, int &holes)
{
    using Ops = ColOps<ColT>;
    const ColT hbit = (ColT)1 << H;
    full = hbit - 1;
    any = 0;
    int sf = 0, sp = 0;
#pragma unroll
    for (int x = 0; x < W; ++x) {
        const ColT c = rec.col(x);
        full &= c;
        any |= c;
        sf += Ops::ffs(c | hbit);
        sp += Ops::popc(c);
    }
    holes = W * (H + 1) - sf - sp;
}

// The lock branch (ref:262-299).
template <typename ColT>
__device__ __forceinline__ void tpe_lock(const TpeRec<ColT> &rec, Piece &pc, const unsigned long long *s_cells, const Params &p,
                                         int W, int H, long long e, int &reward, int &done, int &errbits)
{
    using Ops = ColOps<ColT>;
    {   // _set_piece(True) (ref:263): in-board cells only
        const PieceCols<ColT> q = tpe_piece_cols<ColT>(s_cells + 28, pc.id, pc.rot, pc.x, pc.y, W, H);
        ColT v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = q.X[c] >= 0 ? rec.col(q.X[c]) : 0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (q.X[c] >= 0) rec.set_col(q.X[c], v[c] | q.m[c]);
    }
    // One pass over the columns answers _clear_lines' can_clear (ref:206: a row is full when every column has it),
    // _count_holes and sum(np.any(board, axis=0)) (ref:287,289: rows with any cell = popc of the OR of all columns).
    ColT full, any;
    int holes;
    tpe_scan(rec, W, H, full, any, holes);
    const int k = Ops::popc(full);
    if (k) {  // rare: squeeze the full rows out of every column, top-most first (ref:207-214), then recount
        for (int x = 0; x < W; ++x) {
            ColT c = rec.col(x);
            ColT f = full;
            while (f) {
                const ColT bit = f & (~f + 1);  // lowest set bit = top-most full row r
                f ^= bit;
                c = (c & ~(bit | (bit - 1))) | ((c & (bit - 1)) << 1);  // rows below r stay, rows above move down by one
            }
            rec.set_col(x, c);
        }
        ColT dummy;
        tpe_scan(rec, W, H, dummy, any, holes);
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
__device__ __forceinline__ void tpe_engine_step(const TpeRec<ColT> &rec, int action, const Params &p, int W, int H, long long e,
                                                const unsigned long long *s_cells, int &reward, int &done, int &errbits)
{
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
        if (ld >= p.lock_mod) {                          // ref:258 `% (lock_delay + 1)`: one subtraction unless a
            ld -= p.lock_mod;                            // counter beyond the modulus was injected
            if (ld >= p.lock_mod) ld %= p.lock_mod;
        }
        if (ld == 0) tpe_lock(rec, pc, s_cells, p, W, H, e, reward, done, errbits);
    }
    rec.w[1] = (uint32_t)ld;
    rec.w[0] = (uint32_t)pack_piece(pc);
}

__device__ __forceinline__ uint32_t smem_u32(const void *ptr) { return (uint32_t)__cvta_generic_to_shared(ptr); }

// L2 eviction priorities (Params::tpe_l2 bit 0: observations / info are write-once streams -> evict_first;
// bit 1: env records are re-read by the next step -> evict_last)
__device__ __forceinline__ unsigned long long l2_policy(int kind)  // 0 normal, 1 evict_first, 2 evict_last
{
    unsigned long long pol;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
#ifndef ST_TPE_PLAIN_ST
#define ST_TPE_PLAIN_ST 0  // experiment knob: 1 = ordinary stores without L2 cache hints
#endif
#ifndef ST_TPE_NO_CPASYNC
#define ST_TPE_NO_CPASYNC 0  // experiment knob: 1 = records through registers (LDG + STS) instead of cp.async
#endif
#if ST_TPE_PLAIN_ST
__device__ __forceinline__ void stg128(void *ptr, const float4 &v, unsigned long long) { *reinterpret_cast<float4 *>(ptr) = v; }
__device__ __forceinline__ void stg128(void *ptr, const uint4 &v, unsigned long long) { *reinterpret_cast<uint4 *>(ptr) = v; }
__device__ __forceinline__ void stg32(void *ptr, uint32_t v, unsigned long long) { *reinterpret_cast<uint32_t *>(ptr) = v; }
#else
__device__ __forceinline__ void stg128(void *ptr, const float4 &v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol));
}
__device__ __forceinline__ void stg128(void *ptr, const uint4 &v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "l"(pol));
}
__device__ __forceinline__ void stg32(void *ptr, uint32_t v, unsigned long long pol)
{
    asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(ptr), "r"(v), "l"(pol));
}
#endif

// nibble << 4 -> four float32 cells (ref:400)
__device__ __forceinline__ float4 tpe_slot(const float4 *s_lut, uint32_t off)
{
    return *reinterpret_cast<const float4 *>(reinterpret_cast<const char *>(s_lut) + off);
}

// Records of one group: HBM -> shared memory, asynchronously (one cp.async group per call, possibly empty).
__device__ __forceinline__ void tpe_fetch(uint32_t *recs, const unsigned char *state, int e0, int nvalid, int SW, int pitch,
                                          int stride, uint32_t inv_sw20, int lane, unsigned long long pol)
{
    if (nvalid > 0) {
        const uint32_t *g_rec = reinterpret_cast<const uint32_t *>(state + (long long)e0 * stride);
        const int nwords = nvalid * SW;
#if ST_TPE_NO_CPASYNC
        if (pitch == SW && (reinterpret_cast<uintptr_t>(g_rec) & 15) == 0) {
            const int nvec = nwords >> 2;
            for (int i = lane; i < nvec; i += 32) reinterpret_cast<uint4 *>(recs)[i] = reinterpret_cast<const uint4 *>(g_rec)[i];
            for (int i = (nvec << 2) + lane; i < nwords; i += 32) recs[i] = g_rec[i];
        } else
#endif
        if (pitch == SW && (reinterpret_cast<uintptr_t>(g_rec) & 15) == 0) {  // contiguous in both, 16-byte aligned
            const int nvec = nwords >> 2;
            for (int i = lane; i < nvec; i += 32)
                asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(recs + 4 * i)), "l"(g_rec + 4 * i), "l"(pol) : "memory");
            for (int i = (nvec << 2) + lane; i < nwords; i += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(recs + i)), "l"(g_rec + i) : "memory");
        } else {
            for (int i = lane; i < nwords; i += 32) {
                const int r = (int)(((uint32_t)i * inv_sw20) >> 20);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(recs + r * pitch + (i - r * SW))), "l"(g_rec + i) : "memory");
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// Observations of one group by direct 16-byte stores (float32 [W][H] per env, ref:421-424, 400; needs H % 4 == 0 — the
// launcher stages everything else).  Two passes, each with the lane mapping that makes its addressing trivial:
//  a. one lane per (env, column): the column's H / 4 nibbles go into a byte array in slot order — (env, column) pairs
//     and the float4 slots of a column are both consecutive in the group's block — as table OFFSETS;
//  b. one lane per float4 slot: byte -> table entry -> store, 512 contiguous bytes per warp instruction, four
//     independent stores in flight per lane (a store holds its source registers until the LSU has read them).
template <typename ColT, int WCT, int HCT>
__device__ __forceinline__ void tpe_obs_direct(const uint32_t *recs, unsigned char *stage, unsigned char *dst,
                                            const float4 *s_lut, int nvalid, int Wrt, int Hrt, int pitch, uint32_t inv_w20, bool u8,
                                            int lane, unsigned long long pol_out)
{
    constexpr int CW = ColOps<ColT>::kWords;
    const int W = WCT ? WCT : Wrt, H = HCT ? HCT : Hrt;
    const int hq = H >> 2, nq = W * hq, total = nvalid * nq, items = nvalid * W;
    for (int it = lane; it < items; it += 32) {
        const int r = WCT ? (int)((unsigned)it / (unsigned)W) : (int)(((uint32_t)it * inv_w20) >> 20);
        const int x = it - r * W;
        const uint32_t *cw = recs + r * pitch + kStateWords + CW * x;
        unsigned char *d = stage + it * hq;
        if constexpr (CW == 1) {
            const uint32_t c0 = cw[0];
            const uint32_t w = c0 << 4;
#pragma unroll
            for (int k = 0; k < hq; ++k) d[k] = (unsigned char)((w >> (4 * k)) & 0xf0u);
        } else {
            const uint32_t c0 = cw[0], c1 = cw[1];
            const unsigned long long w = ((unsigned long long)c1 << 32) | c0;
#pragma unroll
            for (int k = 0; k < hq; ++k) d[k] = (unsigned char)(((w >> (4 * k)) & 15u) << 4);
        }
    }
    __syncwarp();
    if (!u8) {
        // whole rounds of 4 x 32 slots: no clamps, no predicates, every offset an immediate of one running pointer
        const unsigned char *sp = stage + lane;
        float4 *dp = reinterpret_cast<float4 *>(dst) + lane;
        const int nfull = total & ~127;
        for (int base = 0; base < nfull; base += 128, sp += 128, dp += 128) {
            uint32_t off[4];
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) off[u] = sp[32 * u];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = tpe_slot(s_lut, off[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) stg128(dp + 32 * u, v[u], pol_out);
        }
        if (nfull < total) {  // the last, partial round: clamped loads, predicated stores
            uint32_t off[4];
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) off[u] = stage[min(nfull + lane + 32 * u, total - 1)];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = tpe_slot(s_lut, off[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (nfull + lane + 32 * u < total) stg128(dp + 32 * u, v[u], pol_out);
        }
    } else {
        for (int s4 = lane; s4 < total; s4 += 32)
            stg32(reinterpret_cast<uint32_t *>(dst) + s4, (((uint32_t)stage[s4] >> 4) * 0x00204081u) & 0x01010101u, pol_out);
    }
}

// WCT / HCT: board size known at compile time (0 = taken from Params): the column loops unroll and the divisions by
// W, H / 4 fold into constants for the boards every BASELINE.json workload uses.
template <typename ColT, int WCT, int HCT>
__device__ __forceinline__ void tpe_step_body(const Params &p)
{
    extern __shared__ __align__(128) uint32_t s_dyn[];
    __shared__ unsigned long long s_cells[56];  // CellTab: e[28], then c[28]
    __shared__ __align__(16) float4 s_lut[16];  // four cells -> four float32 (ref:400)
    constexpr int CW = ColOps<ColT>::kWords;
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpc = blockDim.x >> 5;
#if ST_TPE_TRACE
    if (g_tpe_trace && lane == 0 && (size_t)blockIdx.x * wpc + warp < kTraceWarps)
        g_tpe_trace[((size_t)(p.draw_piece & 7) * kTraceWarps + (size_t)blockIdx.x * wpc + warp) * 16 + 9] = tpe_now();
#endif
    for (int i = threadIdx.x; i < 56; i += blockDim.x) s_cells[i] = __ldg(reinterpret_cast<const unsigned long long *>(&g_cells) + i);
    if (threadIdx.x < 16)
        s_lut[threadIdx.x] = make_float4((threadIdx.x & 1) ? 1.0f : 0.0f, (threadIdx.x & 2) ? 1.0f : 0.0f,
                                         (threadIdx.x & 4) ? 1.0f : 0.0f, (threadIdx.x & 8) ? 1.0f : 0.0f);
    const int H = HCT ? HCT : p.H, W = WCT ? WCT : p.W;
    const int SW = (WCT ? kStateWords + WCT * CW : p.stride >> 2);  // words per record in HBM
    const int pitch = SW | 1;      // odd pitch in smem: lane r, word c -> bank (r * pitch + c) % 32, conflict-free
    const int epw = p.tpe_epw;     // envs per group (32, 16, 8 or 4): fewer envs per warp = more warps for mid-size batches
    const bool u8 = p.obs_u8 != 0;
    const int nel = W * H;
    // warp-private shared memory: two record buffers (this group, next group), then the observation staging block
    const int rec_words = (epw * pitch + 3) & ~3;
    const bool staged = p.tpe_staged != 0;
    // staged: two chunk buffers of 32 (env, column) items; direct: one byte per float4 slot of the group's block
    const int chunk_bytes = (32 * H * (u8 ? 1 : 4) + 15) & ~15;
    const int stage_words = staged ? 2 * chunk_bytes >> 2 : (epw * (nel >> 2) + 15) >> 4 << 2;
    const int nrec = p.tpe_nrec;  // record buffers per warp: 2 when warps walk over several groups (next group in flight)
    uint32_t *const wbase = s_dyn + (size_t)warp * (nrec * rec_words + stage_words);
    unsigned char *const stage = reinterpret_cast<unsigned char *>(wbase + nrec * rec_words);
    const int n32 = (int)p.n;  // launch_tpe refuses batches beyond 2^30 envs
    const int epw_log2 = 31 - __clz(epw);
    const int ngroups = (n32 + epw - 1) >> epw_log2;
    const int gstride = (int)gridDim.x * wpc;
    int g = (int)blockIdx.x * wpc + warp;
    __syncthreads();  // s_cells, s_lut
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (g >= ngroups) {
        if (p.tpe_sync) __syncthreads();  // the phase barrier below counts every warp of the CTA
        return;
    }
    TPE_MARK(0);

    int errbits = 0;
    const size_t n_envs = (size_t)p.n;
    const unsigned long long pol_out = l2_policy((p.tpe_l2 & 1) ? 1 : 0), pol_state = l2_policy((p.tpe_l2 & 2) ? 2 : 0);
    if (nrec == 2) {
        const int e0 = g << epw_log2;
        const int nv = min(n32 - e0, epw);
        tpe_fetch(wbase, p.state, e0, nv, SW, pitch, p.stride, p.inv_sw20, lane, pol_state);
    }
    int cur = 0, chunk = 0;
    for (; g < ngroups; g += gstride, cur ^= nrec - 1) {
    uint32_t *const recs = wbase + cur * rec_words;
    const int e0 = g << epw_log2;
    const int nvalid = min(n32 - e0, epw);
    const int e = e0 + lane;
    unsigned int action_u = 6u;  // issued before the wait so that its miss overlaps the record copy
    if (lane < nvalid) asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(action_u) : "l"(p.actions + e));
    if (nrec == 2) {  // 1. next group's records on their way; this group's have arrived
        const int gn = g + gstride;
        const int en = gn << epw_log2;
        const int nvn = gn < ngroups ? min(n32 - en, epw) : 0;
        tpe_fetch(wbase + (cur ^ 1) * rec_words, p.state, en, nvn, SW, pitch, p.stride, p.inv_sw20, lane, pol_state);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
        tpe_fetch(wbase, p.state, e0, nvalid, SW, pitch, p.stride, p.inv_sw20, lane, pol_state);
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncwarp();
    TPE_MARK(1);

    const TpeRec<ColT> rec = {recs + lane * pitch};
    for (int t = 0; t < p.T; ++t) {  // st_step_many: the records stay in shared memory between steps
    // 2. engine, one env per lane
    int reward = 0, done = 0;
    if (lane < nvalid) tpe_engine_step(rec, (int)action_u, p, W, H, e, s_cells, reward, done, errbits);
    if (t + 1 < p.T && lane < nvalid)  // next step's action, in flight during the cooperative phases
        asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(action_u) : "l"(p.actions + (size_t)(t + 1) * n_envs + e));
    __syncwarp();
    TPE_MARK(2);

    if (p.term_obs) {  // terminal observation of the envs that end here: their board already holds the locked piece
        unsigned term = __ballot_sync(FULL, lane < nvalid && done && p.auto_reset);
        char *tb = reinterpret_cast<char *>(p.term_obs) + ((long long)t * p.obs_t_stride + (long long)e0 * p.obs_elems) * (u8 ? 1 : 4);
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
    // 3. what the observation shows: an auto-reset env shows the empty board of clear() (ref:306-315), every other
    //    lane ORs its piece into its columns (_set_piece(True), ref:301).  The observation goes out FIRST (it is 80 % of
    //    the bytes and the store pipe is what the kernel waits for); info is read before the reset touches the counters.
    const bool resets = lane < nvalid && done && p.auto_reset;
    // Few values stay live across the observation loop (its stores want registers of their own): the piece's
    // column masks and the four column numbers packed into one word (0xff = none).
    ColT shown_m[4] = {0, 0, 0, 0};
    uint32_t shown_x = 0xffffffffu;
    if (lane < nvalid) {
        p.reward[(size_t)t * n_envs + e] = (float)reward;
        p.done[(size_t)t * n_envs + e] = (unsigned char)done;
        if (done && p.stats) {  // episode statistics (K4), from the counters as the terminal step left them
            atomicAdd(p.stats, 1ull);
            atomicAdd(p.stats + 1, (unsigned long long)(long long)(int)rec.w[2]);
            atomicAdd(p.stats + 2, (unsigned long long)(long long)(int)rec.w[4]);
            atomicAdd(p.stats + 3, (unsigned long long)(long long)(int)rec.w[3]);
        }
        if (resets) {
            for (int i = 0; i < W * CW; ++i) rec.w[kStateWords + i] = 0u;
        } else {
            const Piece pc = unpack_piece((int)rec.w[0]);
            if (pc.id < 7) {
                const PieceCols<ColT> q = tpe_piece_cols<ColT>(s_cells + 28, pc.id, pc.rot, pc.x, pc.y, W, H);
                ColT under[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) under[c] = q.X[c] >= 0 ? rec.col(q.X[c]) : 0;
                shown_x = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (q.X[c] >= 0) rec.set_col(q.X[c], under[c] | q.m[c]);
                    shown_m[c] = q.m[c];
                    shown_x |= (uint32_t)(q.X[c] & 0xff) << (8 * c);
                }
            }
        }
    }
    TPE_MARK(3);
    __syncwarp();
    // Experiment knob (ST_B200_TPE_SYNC=1, off): ONE CTA per SM and nobody stores before every warp of the SM has stepped
    // its envs.  Stores and shared-memory loads share the LSU queue, so without the barrier the engines of late warps
    // crawl behind the early warps' stores (engine done: p95 7.9 us, max 10.5 us at 65 536 envs; with the barrier 5.0 /
    // 5.7 us).  But once every SM stores at the same time the stream runs at the HBM write rate (59 MB in 7.2 us), and
    // the launch ends later than with the ragged overlap (16.7 against 14.7 us): profiles/r2_tpe_phase_barrier_ab.txt.
    if (p.tpe_sync) __syncthreads();
    TPE_MARK(4);

    // 4. observations: float32 [W][H] per env (ref:421-424, 400).  One lane per (env, column): the column's cells are
    //    consecutive in the observation, (env, column) pairs are consecutive in the group's block.
    if (p.obs && !staged) {
        unsigned char *dst = reinterpret_cast<unsigned char *>(p.obs) +
                             ((long long)t * p.obs_t_stride + (long long)e0 * p.obs_elems) * (u8 ? 1 : 4);
        tpe_obs_direct<ColT, WCT, HCT>(recs, stage, dst, s_lut, nvalid, W, H, pitch, p.inv_w20, u8, lane, pol_out);
        __syncwarp();
    } else if (p.obs) {
        // staged: 32 (env, column) items at a time — H consecutive cells each, and consecutive items are consecutive
        // in the group's block — are expanded into one of two warp-private chunk buffers, and ONE cp.async.bulk (TMA)
        // store per chunk sends it on its way: no store instruction queues in the LSU in front of the other warps'
        // shared-memory loads, and a buffer is only waited for when it comes round again two chunks later.
        const int items = nvalid * W;
        const int item_bytes = H * (u8 ? 1 : 4);
        unsigned char *dstb = reinterpret_cast<unsigned char *>(p.obs) +
                              ((long long)t * p.obs_t_stride + (long long)e0 * p.obs_elems) * (u8 ? 1 : 4);
        for (int it0 = 0; it0 < items; it0 += 32, chunk ^= 1) {
            unsigned char *buf = stage + chunk * chunk_bytes;
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // the store of two chunks ago has read `buf`
            __syncwarp();
            const int it = it0 + lane;
            if (it < items) {
                const int r = WCT ? (int)((unsigned)it / (unsigned)W) : (int)(((uint32_t)it * p.inv_w20) >> 20);
                const int x = it - r * W;
                const uint32_t *cw = recs + r * pitch + kStateWords + CW * x;
                unsigned char *d = buf + lane * item_bytes;
                if ((H & 3) == 0) {
                    const int hq = H >> 2;
                    if (!u8) {
#pragma unroll
                        for (int k = 0; k < hq; ++k)
                            reinterpret_cast<float4 *>(d)[k] = s_lut[(cw[k >> 3] >> (4 * (k & 7))) & 15u];
                    } else {
#pragma unroll
                        for (int k = 0; k < hq; ++k)  // bit b of the nibble -> byte b (the four partial products do not overlap)
                            reinterpret_cast<uint32_t *>(d)[k] = (((cw[k >> 3] >> (4 * (k & 7))) & 15u) * 0x00204081u) & 0x01010101u;
                    }
                } else {
                    for (int y = 0; y < H; ++y) {
                        const bool on = ((cw[y >> 5] >> (y & 31)) & 1u) != 0u;
                        if (u8) d[y] = on ? 1 : 0;
                        else reinterpret_cast<float *>(d)[y] = on ? 1.0f : 0.0f;
                    }
                }
            }
            const uint32_t bytes = (uint32_t)((items - it0 < 32 ? items - it0 : 32) * item_bytes);
            unsigned char *dst = dstb + (size_t)it0 * item_bytes;
            const bool bulk = ((reinterpret_cast<uintptr_t>(dst) | bytes) & 15) == 0;
            if (bulk) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // every lane: its chunk writes, for the TMA engine
            __syncwarp();
            if (!bulk) {  // ragged tails of uint8 / odd-height batches: plain stores from the chunk buffer
                if (((reinterpret_cast<uintptr_t>(dst) | bytes) & 3) == 0)
                    for (uint32_t i = lane; i < (bytes >> 2); i += 32) reinterpret_cast<uint32_t *>(dst)[i] = reinterpret_cast<const uint32_t *>(buf)[i];
                else
                    for (uint32_t i = lane; i < bytes; i += 32) dst[i] = buf[i];
                __syncwarp();
            }
            if (lane == 0) {
                if (bulk) asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(buf)), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");  // one group per chunk, empty or not: the wait above counts groups
            }
        }
    } else {
        __syncwarp();
    }

    TPE_MARK(5);
    // 5. info: the counters as the step left them, before any reset (reward and done went out before the observation)
    if (p.info) {
        int32_t *g_info = p.info + (long long)t * p.info_t_stride + (long long)e0 * kStateWords;
        const int ninfo = nvalid * kStateWords;
        if (((reinterpret_cast<uintptr_t>(g_info) & 15) | (nvalid & 3)) == 0) {
            // 16-byte stores: the four words of slot q are words 4q..4q+3 of the group's [nvalid][15] block, i.e. word
            // c, c+1, .. of record r = 4q / 15, running over into record r + 1 past word 14
            for (int q = lane; q < (ninfo >> 2); q += 32) {
                const int r = (q * 4 * 4370) >> 16;  // 4q / 15 for 4q < 480
                const int c = 4 * q - r * kStateWords;
                const uint32_t *src = recs + r * pitch + c;
                uint32_t v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = src[c + k >= kStateWords ? k + pitch - kStateWords : k];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = (c + k == 0 || c + k == kStateWords) ? (v[k] & 15u) : v[k];  // word 0: the piece id only
                stg128(reinterpret_cast<uint4 *>(g_info) + q, make_uint4(v[0], v[1], v[2], v[3]), pol_out);
            }
        } else {
            for (int j = lane; j < ninfo; j += 32) {
                const int r = (j * 4370) >> 16;  // j / 15 for j < 480
                const int c = j - r * kStateWords;
                const uint32_t v = recs[r * pitch + c];
                stg32(g_info + j, c == 0 ? (v & 15u) : v, pol_out);
            }
        }
    }
    __syncwarp();
    // 6. _set_piece(False) (ref:303), literally: the cells of the piece are cleared on the board; an auto-reset env
    //    gets the rest of clear(): zeroed episode counters and a fresh piece
    if (lane < nvalid) {
        ColT shown_c[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int X = (int)((shown_x >> (8 * c)) & 0xffu);
            shown_c[c] = X != 0xff ? rec.col(X) : 0;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int X = (int)((shown_x >> (8 * c)) & 0xffu);
            if (X != 0xff) rec.set_col(X, shown_c[c] & ~shown_m[c]);
        }
        if (resets) {
#pragma unroll
            for (int i = 2; i <= 6; ++i) rec.w[i] = 0u;
            Piece pc;
            pc.id = tpe_spawn(rec, p, e, errbits);
            pc.rot = 0; pc.x = W / 2; pc.y = 0;
            rec.w[0] = (uint32_t)pack_piece(pc);
        }
    }
    __syncwarp();
    }  // for t
    {   // records back to HBM
        uint32_t *g_rec = reinterpret_cast<uint32_t *>(p.state + (long long)e0 * p.stride);
        const int nwords = nvalid * SW;
        if (pitch == SW && (reinterpret_cast<uintptr_t>(g_rec) & 15) == 0) {
            const int nvec = nwords >> 2;
#pragma unroll 4
            for (int i = lane; i < nvec; i += 32) stg128(reinterpret_cast<uint4 *>(g_rec) + i, reinterpret_cast<const uint4 *>(recs)[i], pol_state);
            for (int i = (nvec << 2) + lane; i < nwords; i += 32) g_rec[i] = recs[i];
        } else {
            for (int i = lane; i < nwords; i += 32) {
                const int r = (int)(((uint32_t)i * p.inv_sw20) >> 20);
                g_rec[i] = recs[r * pitch + (i - r * SW)];
            }
        }
    }
    __syncwarp();
    TPE_MARK(6);
    }  // for g
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (staged && lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // smem must outlive the TMA reads
#if ST_TPE_TRACE
    if (g_tpe_trace && lane == 0 && (size_t)blockIdx.x * wpc + warp < kTraceWarps) {
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        g_tpe_trace[((size_t)(p.draw_piece & 7) * kTraceWarps + (size_t)blockIdx.x * wpc + warp) * 16 + 7] = tpe_now();
        g_tpe_trace[((size_t)(p.draw_piece & 7) * kTraceWarps + (size_t)blockIdx.x * wpc + warp) * 16 + 8] = smid;
    }
#endif
    errbits = __reduce_or_sync(FULL, errbits);
    if (errbits && p.err && lane == 0) atomicOr(p.err, errbits);
}

// Two entry points of the same body.  The default build is capped at 64 registers (32 resident warps per SM: what
// multi-wave batches, T-step launches and the 64-bit-column boards want).  The second one gets 72 registers — 7 CTAs of
// 4 warps per SM, i.e. the 28 warps per SM of a one-wave batch with nothing to spare: measured on B200, one-step
// launches of 10x20 boards, us per launch, 64 / 72 registers: 16384 envs 8.76 / 7.92, 32768 envs 10.18 / 9.50,
// 65536 envs 14.61 / 14.01, 262144 envs 44.0 / 43.7; 20x40 boards: 16384 envs 13.9 / 13.0, 32768 envs 23.0 / 22.2, but
// 65536 envs 39.5 / 40.5, and T = 32: 9.46 / 9.56 — those keep 64.
template <typename ColT, int WCT, int HCT>
__global__ void __launch_bounds__(32 * kTpeMaxWarps, ST_TPE_MINBLOCKS) st_step_tpe_kernel(const __grid_constant__ Params p)
{
    tpe_step_body<ColT, WCT, HCT>(p);
}
#ifndef ST_TPE_R72_NREG
#define ST_TPE_R72_NREG 72  // experiment knob
#endif
template <typename ColT, int WCT, int HCT>
__global__ void __maxnreg__(ST_TPE_R72_NREG) st_step_tpe_kernel_r72(const __grid_constant__ Params p)
{
    tpe_step_body<ColT, WCT, HCT>(p);
}

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// Launch shape of the thread-per-env kernel: envs per group, warps per CTA, grid cap (CTAs per SM; 0 = one group per warp).
struct TpeShape {
    int epw, wpc, ctas_per_sm, staged, sync;
};

static int tpe_sm_count()
{
    static int n_sm[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (!n_sm[dev]) cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
    return n_sm[dev] > 0 ? n_sm[dev] : 148;
}

static size_t tpe_smem_bytes(const Params &p, int epw, int wpc, int staged, int nrec)
{
    const int pitch = (p.stride >> 2) | 1;
    const int rec_words = (epw * pitch + 3) & ~3;
    const int chunk_bytes = (32 * p.H * (p.obs_u8 ? 1 : 4) + 15) & ~15;
    const int stage_words = staged ? 2 * chunk_bytes >> 2 : (epw * (p.W * p.H >> 2) + 15) >> 4 << 2;
    return (size_t)wpc * (nrec * rec_words + stage_words) * 4;
}

constexpr size_t kTpeSmemMax = 227 * 1024 - 1024;  // per CTA, minus the static tables and the per-CTA reserve

template <typename ColT, int WCT, int HCT>
static cudaError_t launch_tpe_t(const Params &p, const TpeShape &s, cudaStream_t stream)
{
    const long long ngroups = (p.n + s.epw - 1) / s.epw;
    long long nctas = (ngroups + s.wpc - 1) / s.wpc;
    // ST_B200_TPE_SMEM_PAD: experiment knob, unused dynamic shared memory that caps the resident CTAs per SM (waves)
    size_t smem = tpe_smem_bytes(p, s.epw, s.wpc, s.staged, p.tpe_nrec) + (size_t)env_int("ST_B200_TPE_SMEM_PAD", 0);
    if (smem > kTpeSmemMax) smem = kTpeSmemMax;
    static const bool pdl = getenv("ST_B200_NO_PDL") == nullptr;
    int dev = 0;
    cudaGetDevice(&dev);
    // 72-register entry point: one-step launches on 32-bit columns, CTAs of at most 28 warps (ST_B200_TPE_R72 = 0 / 1 forces)
    const int r72_knob = env_int("ST_B200_TPE_R72", -1);
    const bool r72 = s.wpc <= 28 && (r72_knob >= 0 ? r72_knob != 0 : (p.T == 1 && (sizeof(ColT) == 4 || p.n < 49152)));
    const auto kernel = r72 ? st_step_tpe_kernel_r72<ColT, WCT, HCT> : st_step_tpe_kernel<ColT, WCT, HCT>;
    static bool attr_set[64] = {};
    static int n_sm[64] = {};
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(st_step_tpe_kernel<ColT, WCT, HCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTpeSmemMax);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(st_step_tpe_kernel_r72<ColT, WCT, HCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTpeSmemMax);
        if (e != cudaSuccess) return e;
        cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
        attr_set[dev] = true;
    }
    if (s.ctas_per_sm > 0 && dev >= 0 && dev < 64 && n_sm[dev] > 0) {
        const long long cap = (long long)s.ctas_per_sm * n_sm[dev];
        nctas = nctas < cap ? nctas : cap;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)nctas);
    cfg.blockDim = dim3(32 * s.wpc);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    count_launch();
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

// Which batches take this kernel and with how many envs per group: measured on B200 (tools/knob_sweep.py; us per
// one-step launch, 10x20 boards, 72-register entry point; cols = the column-lane warp-per-env kernel):
//   n         cols    epw 4    epw 8    epw 16         n         epw 4    epw 8    epw 16
//   8192      6.18     5.98     7.97     9.27         32768      14.11     9.55    11.81
//   10240     7.22     7.71     8.35     9.43         40960      16.92    12.33    13.10
//   12288     8.23     7.83     8.71    10.22         49152      20.02    13.52    13.43
//   14336     9.22     7.97     8.61    10.69         57344      22.92    15.03    14.57
//   16384    ~10.3     7.93     7.50    10.45         65536      25.56    16.20    14.01
//   24576       -     11.26     9.17    11.91        131072      49.54    30.82    24.82
// The per-warp engine chain costs the same for 4 or 32 envs, so mid-size batches get fewer envs per warp (more warps
// to hide its latency) and large ones more (fewer instructions per env); a group count just under one wave (4144 warps
// at 7 CTAs per SM) is the sweet spot, a little over it the worst case.
// 20 wide x 40 high boards: 4 envs per group below 24576 envs (16384 envs: 13.0 against 15.2 us), 8 above.
// T steps per launch (records stay in shared memory between steps): 8 envs per group.
// Below the thresholds the column-lane warp-per-env kernel (st_kernels_cols.cuh) is as fast or faster (table above; at
// T = 32, us per step, column lanes / thread-per-env: 6144 envs 2.3 / 2.6, 8192 2.8 / 2.75, 12288 4.0 / 3.1).
static long long tpe_min_envs(const Params &p) { return p.T > 1 ? 8192 : 11264; }
static int tpe_default_epw(const Params &p)
{
    if (p.T > 1) return 8;
    if (p.H <= 31) return p.n < 16384 ? 4 : p.n < 49152 ? 8 : 16;
    return p.n < 24576 ? 4 : 8;
}

static TpeShape tpe_shape(const Params &p)
{
    TpeShape s;
    s.epw = env_int("ST_B200_TPE_EPW", tpe_default_epw(p));
    if (s.epw != 4 && s.epw != 8 && s.epw != 16 && s.epw != 32) s.epw = 32;
    s.wpc = env_int("ST_B200_TPE_WPC", 4);
    if (s.wpc < 1 || s.wpc > kTpeMaxWarps) s.wpc = 4;
    // Experiment (off): single-step launches whose groups fit the machine in one wave as ONE CTA per SM with a barrier
    // between the engine and the store phases (see the kernel; measured slower, profiles/r2_tpe_phase_barrier_ab.txt).
    s.sync = 0;
    {
        const long long ngroups = (p.n + s.epw - 1) / s.epw;
        const int n_sm = tpe_sm_count();
        const long long per_sm = (ngroups + n_sm - 1) / n_sm;
        const int want = env_int("ST_B200_TPE_SYNC", 0);
        if (want && p.T == 1 && per_sm <= kTpeMaxWarps && per_sm >= env_int("ST_B200_TPE_SYNC_MIN", 8) && getenv("ST_B200_TPE_WPC") == nullptr &&
            getenv("ST_B200_TPE_CTAS_PER_SM") == nullptr) {
            s.wpc = (int)per_sm;
            s.sync = 1;
        } else if (want == 2 && p.T == 1) {
            s.sync = 1;  // experiment: barrier with whatever CTA shape was asked for
        }
    }
    // boards whose columns are not whole float4s always go through the staging block
    s.staged = (p.H & 3) != 0 ? 1 : env_int("ST_B200_TPE_STAGED", 0);
    s.ctas_per_sm = env_int("ST_B200_TPE_CTAS_PER_SM", 0);
    const int nrec = s.ctas_per_sm > 0 ? 2 : 1;
    if (tpe_smem_bytes(p, s.epw, s.wpc, s.staged, nrec) > kTpeSmemMax) s.sync = 0;
    while (s.wpc > 1 && tpe_smem_bytes(p, s.epw, s.wpc, s.staged, nrec) > kTpeSmemMax) s.wpc >>= 1;
    while (s.epw > 4 && tpe_smem_bytes(p, s.epw, s.wpc, s.staged, nrec) > kTpeSmemMax) s.epw >>= 1;
    return s;
}

// Thread-per-env path: ram observations, single-step launches and st_step_many alike.
static bool tpe_eligible(const Params &p, int obs_type)
{
    return obs_type == 0 && p.mode == MODE_STEP && p.n > 0 && p.n <= (1ll << 30) && (p.obs_t_stride & 3) == 0 &&
           tpe_smem_bytes(p, 4, 1, 1, 2) <= kTpeSmemMax;
}

static cudaError_t launch_tpe(const Params &p0, cudaStream_t stream)
{
    Params p = p0;
    const TpeShape s = tpe_shape(p);
    p.tpe_epw = s.epw;
    p.tpe_staged = s.staged;
    p.tpe_sync = s.sync;
    p.tpe_nrec = s.ctas_per_sm > 0 ? 2 : 1;
    p.tpe_l2 = env_int("ST_B200_TPE_L2", 1);  // observations / info leave as evict_first streams (measured: -3..5 %)
#if ST_TPE_TRACE
    static int launch_id = 0;
    p.draw_piece = launch_id++;  // unused by step launches: which of the 8 trace slabs this launch writes
#endif
    if (p.col_words == 1) {
        if (p.W == 10 && p.H == 20) return launch_tpe_t<uint32_t, 10, 20>(p, s, stream);
        return launch_tpe_t<uint32_t, 0, 0>(p, s, stream);
    }
    if (p.W == 20 && p.H == 40) return launch_tpe_t<unsigned long long, 20, 40>(p, s, stream);
    return launch_tpe_t<unsigned long long, 0, 0>(p, s, stream);
}

}  // namespace st

#if ST_TPE_TRACE
// profiling builds only (never part of the product library): where the kernel writes its phase timestamps
extern "C" __attribute__((visibility("default"))) int st_debug_set_tpe_trace(unsigned long long *buf)
{
    return (int)cudaMemcpyToSymbol(st::g_tpe_trace, &buf, sizeof(buf));
}
#endif
