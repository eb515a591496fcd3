// SimpleTetris step path for sm_100a: one warp per env, one board row per lane.
//
// What the reference does in tetris_env.py:10-335 + 397-433 (cited as ref:LINE) on a dense float64
// (W,H) array is done here on a row bitboard held in registers: lane l of the env's warp owns row l
// (and row l+32 when H > 31).  A row register is "widened": board column x sits at bit x+4, bits 0..3
// and every bit from W+4 up are permanently-set WALL bits.  A piece row mask shifted to column x is
// then tested with one AND per row — walls need no separate test, and rows above the board have no lane,
// so they are skipped for board AND walls exactly like ref:32-33.
// Four ballots (one per piece row) give the collision answer for EVERY anchor height at once, so soft
// drop, hard drop (one ctz), gravity and the grounded test all read the same mask.  Line clear is
// compare-to-all-ones + ballot + shuffle compaction, holes are a shuffle prefix-OR + popc, height is popc
// of a ballot.  Control flow is uniform per warp (one env), so the lock / spawn / reset branches cost
// nothing when they are not taken.
// Observations: ram is expanded by the env's own warp from a smem staging row with 16-byte stores;
// 84x84 images are written by the whole CTA (8 envs) with every thread owning a fixed 16-byte column
// slot, so a warp store covers 512 contiguous bytes; rgb leaves through TMA bulk stores from a shared-memory ring.
// Ram batches are stepped by two other kernels since round 2 (see ram_path / launch_main): small ones by the
// column-lane warp-per-env kernel of st_kernels_cols.cuh, large ones by the thread-per-env kernel of st_kernels_tpe.cuh;
// this kernel keeps the image modes, reset / observe launches and ram boards wider than 24 columns.
//
// Compile-time tuning knobs (defaults are the measured best on B200; DESIGN.md section 6 lists what was tried).
#include "st_internal.h"

#include <cstdlib>

namespace st {

#ifndef ST_FORCE_MANY
#define ST_FORCE_MANY 0
#endif
#ifndef ST_ANCHOR_PTRS
#define ST_ANCHOR_PTRS 0
#endif
#ifndef ST_RAM_MINBLOCKS
#define ST_RAM_MINBLOCKS 1
#endif
#ifndef ST_IMG_MINBLOCKS
#define ST_IMG_MINBLOCKS 1
#endif
#ifndef ST_IMG_UNROLL
#define ST_IMG_UNROLL 1
#endif
#ifndef ST_IMG_BULK
#define ST_IMG_BULK 1
#endif
#ifndef ST_IMG_BULK_GRAY
#define ST_IMG_BULK_GRAY 0
#endif
#ifndef ST_IMG_BULK_PASSES
#define ST_IMG_BULK_PASSES 2
#endif

constexpr int kImgUnroll = ST_IMG_UNROLL;  // image-row loop unroll (tuning knob)
constexpr unsigned FULL = 0xffffffffu;
constexpr int OFF = 4;  // board column x lives at bit x + OFF of a widened row

// ---------------------------------------------------------------------------------------------
// Piece table: 7 pieces x 4 rotations (ref:10-19; rotation r = r applications of (i,j)->(j,-i), ref:22-26).
// Entry: bits 0..27 = four 7-bit row masks (row t is piece row j = minj + t; bit i+3 = cell offset i),
//        bits 32..35 = minj + 3, bits 36..39 = maxj + 3.
// ---------------------------------------------------------------------------------------------
struct PieceTab {
    unsigned long long e[28];
};

constexpr PieceTab make_piece_tab()
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
    PieceTab t{};
    for (int id = 0; id < 7; ++id) {
        int c[4][2] = {};
        for (int k = 0; k < 4; ++k) { c[k][0] = base[id][k][0]; c[k][1] = base[id][k][1]; }
        for (int r = 0; r < 4; ++r) {
            int mn = 0, mx = 0;
            for (int k = 0; k < 4; ++k) {
                mn = c[k][1] < mn ? c[k][1] : mn;
                mx = c[k][1] > mx ? c[k][1] : mx;
            }
            unsigned long long m = 0;
            for (int k = 0; k < 4; ++k) m |= 1ull << ((c[k][1] - mn) * 7 + (c[k][0] + 3));
            m |= (unsigned long long)(mn + 3) << 32;
            m |= (unsigned long long)(mx + 3) << 36;
            t.e[id * 4 + r] = m;
            for (int k = 0; k < 4; ++k) { int i = c[k][0], j = c[k][1]; c[k][0] = j; c[k][1] = -i; }
        }
    }
    return t;
}

__constant__ PieceTab c_tab = make_piece_tab();

// Piece rows at column x, as widened-row masks (cell X at bit X + OFF; x + 1 >= 0 since x >= -1).
template <typename RowT>
struct PieceRows {
    RowT m[4];
    int minj, maxj;
};

template <typename RowT>
__device__ __forceinline__ PieceRows<RowT> piece_rows(int id, int rot, int x)
{
    const unsigned long long e = c_tab.e[id * 4 + rot];
    const uint32_t lo = (uint32_t)e, hi = (uint32_t)(e >> 32);
    PieceRows<RowT> pr;
    pr.minj = (int)(hi & 15u) - 3;
    pr.maxj = (int)((hi >> 4) & 15u) - 3;
#pragma unroll
    for (int t = 0; t < 4; ++t) pr.m[t] = (RowT)((lo >> (7 * t)) & 127u) << (x + 1);
    return pr;
}

template <int RPL> struct CMask { using type = uint32_t; };
template <> struct CMask<2> { using type = unsigned long long; };

__device__ __forceinline__ int ctz(uint32_t v) { return __ffs((int)v) - 1; }
__device__ __forceinline__ int ctz(unsigned long long v) { return __ffsll((long long)v) - 1; }

// Collision mask over anchor heights: bit y' set <=> is_occupied(shape, (x, y'), board) (ref:29-36).
template <int RPL, typename RowT>
__device__ __forceinline__ typename CMask<RPL>::type collision_mask(const RowT (&row)[RPL], const PieceRows<RowT> &pr, int H)
{
    using M = typename CMask<RPL>::type;
    M c = 0;
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const M b = __ballot_sync(FULL, (pr.m[t] & row[k]) != 0);  // bit l: board row l+32k meets piece row t
            const int sh = pr.minj + t - 32 * k;                         // anchor y' = row - (minj + t)
            c |= sh >= 0 ? (b >> sh) : (b << -sh);
        }
    }
    int fl = H - pr.maxj;  // anchors whose lowest cell is at or below the floor (ref:34 `y >= board.shape[1]`)
    fl = fl < 0 ? 0 : fl;
    c |= ~(M)0 << fl;
    return c;
}

// Philox4x32-10 keyed by the run seed, counter = (global env id, lifetime piece index).
__device__ __forceinline__ uint32_t philox_draw(uint32_t k0, uint32_t k1, unsigned long long env_id, uint32_t index)
{
    uint32_t c0 = (uint32_t)env_id, c1 = (uint32_t)(env_id >> 32), c2 = index, c3 = 0;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = h1 ^ c1 ^ k0, n2 = h0 ^ c3 ^ k1;
        c0 = n0; c1 = l1; c2 = n2; c3 = l0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c0;
}

// _new_piece / _choose_shape (ref:183-200).  `sw` is this lane's word of the env record (lanes 8..14
// hold shape_counts).  Returns the piece id; bumps its count.
__device__ __forceinline__ int spawn_piece(int &sw, int lane, const Params &p, int e, int &errbits)
{
    int c[7];
    int total = 0, mx = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        c[i] = __shfl_sync(FULL, sw, 8 + i);
        total += c[i];
        mx = c[i] > mx ? c[i] : mx;
    }
    int id;
    if (p.queue) {
        int k = total;
        if (k >= p.queue_len) { errbits |= 1; k %= p.queue_len; }
        id = p.queue[(size_t)e * (unsigned)p.queue_len + k] % 7;
    } else {
        const int S = 35 + 7 * mx - total;  // sum of m_i = 5 + max - c_i (ref:186)
        const uint32_t u = philox_draw(p.seed_lo, p.seed_hi, (unsigned long long)(p.env_id_base + e), (uint32_t)total);
        const int r = 1 + (int)__umulhi(u, (uint32_t)S);  // uniform on [1, S] (ref:187)
        int acc = 0;
        id = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {  // smallest i with cumsum(m)[i] >= r (ref:188-191)
            acc += 5 + mx - c[i];
            id += (r > acc) ? 1 : 0;
        }
    }
    if (lane == 8 + id) sw += 1;
    return id;
}

// Rows [0, fr] shift down by one (row 0 becomes empty): removal of full row `fr` (ref:205-216).
template <int RPL, typename RowT>
__device__ __forceinline__ void remove_row(RowT (&row)[RPL], int fr, int lane, RowT walls)
{
    const RowT up0 = __shfl_up_sync(FULL, row[0], 1);
    if (RPL == 2) {
        const RowT up1 = __shfl_up_sync(FULL, row[RPL - 1], 1);
        const RowT last0 = __shfl_sync(FULL, row[0], 31);
        if (lane + 32 <= fr) row[RPL - 1] = lane ? up1 : last0;
    }
    if (lane <= fr) row[0] = lane ? up0 : walls;
}

// _count_holes (ref:218-220): empty cells with a filled cell above them in the same column.  Wall bits are
// set in every row, so `above & ~row` is zero on them without any masking.
template <int RPL, typename RowT>
__device__ __forceinline__ int count_holes(const RowT (&row)[RPL], int H, int lane)
{
    int cnt = 0;
    RowT carry = 0;  // OR of all rows of the previous 32-row block
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
        RowT v = row[k];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const RowT t = __shfl_up_sync(FULL, v, d);
            if (lane >= d) v |= t;
        }
        RowT above = __shfl_up_sync(FULL, v, 1);
        if (lane == 0) above = 0;
        above |= carry;
        if (lane + 32 * k < H) cnt += sizeof(RowT) == 8 ? __popcll(above & ~row[k]) : __popc((uint32_t)(above & ~row[k]));
        if (k + 1 < RPL) carry |= __shfl_sync(FULL, v, 31);
    }
    return __reduce_add_sync(FULL, cnt);
}

// The piece's cells on this lane's rows (ref:323-327 _set_piece; out-of-board cells land on wall bits).
template <int RPL, typename RowT>
__device__ __forceinline__ void piece_on_rows(RowT (&pm)[RPL], const PieceRows<RowT> &pr, int y, int lane)
{
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
        const int t = lane + 32 * k - y - pr.minj;
        RowT m = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) m = (t == q) ? pr.m[q] : m;
        pm[k] = m;
    }
}

// Env records keep the board as COLUMN words (bit y of column x = cell (x, y); st_kernels_tpe.cuh works on them
// directly).  This kernel wants one ROW per lane: lane x holds column x, every lane gathers its row's bit of each
// column with one shuffle per column, and the other way round with one ballot per column.  Column bits at y >= H
// are zero, so lanes beyond the board come out as bare walls.
template <int RPL, typename RowT>
__device__ __forceinline__ void rows_from_columns(RowT (&row)[RPL], uint32_t clo, uint32_t chi, int W, int lane, RowT walls)
{
    RowT r0 = 0, r1 = 0;
#pragma unroll 4
    for (int x = 0; x < W; ++x) {
        const uint32_t c = __shfl_sync(FULL, clo, x);
        r0 |= (RowT)((c >> lane) & 1u) << x;
        if (RPL == 2) {
            const uint32_t d = __shfl_sync(FULL, chi, x);
            r1 |= (RowT)((d >> lane) & 1u) << x;
        }
    }
    row[0] = (r0 << OFF) | walls;
    if (RPL == 2) row[RPL - 1] = (r1 << OFF) | walls;
}

template <int RPL, typename RowT>
__device__ __forceinline__ void columns_from_rows(const RowT (&row)[RPL], uint32_t &clo, uint32_t &chi, int W, int lane)
{
    clo = 0; chi = 0;
#pragma unroll 4
    for (int x = 0; x < W; ++x) {
        const uint32_t b0 = __ballot_sync(FULL, ((row[0] >> (x + OFF)) & 1) != 0);
        if (lane == x) clo = b0;
        if (RPL == 2) {
            const uint32_t b1 = __ballot_sync(FULL, ((row[RPL - 1] >> (x + OFF)) & 1) != 0);
            if (lane == x) chi = b1;
        }
    }
}

struct Piece {
    int id, rot, x, y;
};
__device__ __forceinline__ int pack_piece(const Piece &pc) { return pc.id | (pc.rot << 4) | (pc.x << 8) | (pc.y << 16); }
__device__ __forceinline__ Piece unpack_piece(int w)
{
    Piece pc;
    pc.id = w & 15; pc.rot = (w >> 4) & 3; pc.x = (w >> 8) & 255; pc.y = (w >> 16) & 255;
    return pc;
}

__device__ __forceinline__ void put(int &sw, int lane, int idx, int v) { if (lane == idx) sw = v; }
__device__ __forceinline__ int get(int sw, int idx) { return __shfl_sync(FULL, sw, idx); }

// clear() (ref:306-315): zero the per-episode counters, spawn, empty board.  The lock-delay counter,
// deaths and shape_counts persist.
template <int RPL, typename RowT>
__device__ __forceinline__ void engine_clear(RowT (&row)[RPL], int &sw, Piece &pc, int lane, const Params &p,
                                             int e, int &errbits, RowT walls)
{
    if (lane >= 2 && lane <= 6) sw = 0;  // time, score, lines_cleared, holes, piece_height
    pc.id = spawn_piece(sw, lane, p, e, errbits);
    pc.rot = 0; pc.x = p.W / 2; pc.y = 0;
#pragma unroll
    for (int k = 0; k < RPL; ++k) row[k] = walls;
}

// The lock branch of TetrisEngine.step (ref:262-299), out of line: taken on ~1/5 of the steps.
template <int RPL, typename RowT>
__device__ __forceinline__ void engine_lock(RowT (&row)[RPL], int &sw, Piece &pc, PieceRows<RowT> &pr, int lane,
                                         const Params &p, int e, int &reward, int &done, int &errbits, RowT walls)
{
    const int H = p.H;
    RowT pm[RPL];
    piece_on_rows<RPL, RowT>(pm, pr, pc.y, lane);
#pragma unroll
    for (int k = 0; k < RPL; ++k) row[k] |= pm[k];  // _set_piece(True) ref:263
    // _clear_lines (ref:205-216): a full row has every bit set (walls included)
    unsigned long long full = 0;
#pragma unroll
    for (int k = 0; k < RPL; ++k)
        full |= (unsigned long long)__ballot_sync(FULL, (lane + 32 * k < H) && row[k] == ~(RowT)0) << (32 * k);
    const int k_cleared = __popcll(full);
    if (k_cleared) {
        unsigned long long f = full;
        while (f) {  // top-most full row first; rows below it keep their index
            const int fr = __ffsll((long long)f) - 1;
            f &= f - 1;
            remove_row<RPL, RowT>(row, fr, lane, walls);
        }
        if (lane == 4) sw += k_cleared;  // lines_cleared (ref:213)
    }
    // line-clear reward / score (ref:266-275)
    int dscore;
    if (p.adv_clears) {
        const int kk = k_cleared > 4 ? 4 : k_cleared;
        dscore = kk == 0 ? 0 : kk == 1 ? 40 : kk == 2 ? 100 : kk == 3 ? 300 : 1200;
        reward += (dscore * 5) / 2;  // 2.5 * {0,40,100,300,1200} is integral
    } else if (p.high_scoring) {
        dscore = k_cleared;
        reward += 1000 * k_cleared;
    } else {
        dscore = k_cleared;
        reward += 100 * k_cleared;
    }
    if (lane == 3) sw += dscore;
    const int old_holes = get(sw, 5);
    const int holes = count_holes<RPL, RowT>(row, H, lane);  // ref:278,284
    put(sw, lane, 5, holes);
    int nonempty = 0;
#pragma unroll
    for (int k = 0; k < RPL; ++k) nonempty += __popc(__ballot_sync(FULL, row[k] != walls));
    const bool top = (__ballot_sync(FULL, row[0] != walls) & 1u) != 0u;  // np.any(board[:,0]) ref:277
    if (top) {
        if (lane == 7) sw += 1;  // n_deaths (ref:279)
        done = 1;
        reward = -100;  // ref:281 overrides everything
    } else {
        if (p.pen_height) {  // ref:286-287: sum(np.any(board, axis=0)) = number of non-empty rows
            reward -= nonempty;
        } else if (p.pen_height_inc) {  // ref:288-292
            const int ph = get(sw, 6);
            if (nonempty > ph) reward -= 10 * (nonempty - ph);
            put(sw, lane, 6, nonempty);
        }
        if (p.pen_holes) reward -= 5 * holes;  // ref:294-297
        else if (p.pen_holes_inc) reward -= 5 * (holes - old_holes);
        pc.id = spawn_piece(sw, lane, p, e, errbits);  // _new_piece ref:299
        pc.rot = 0; pc.x = p.W / 2; pc.y = 0;
        pr = piece_rows<RowT>(pc.id, 0, pc.x);
    }
}

// TetrisEngine.step (ref:243-304).  Outputs reward/done and the display rows (board | piece, ref:301-302).
template <int RPL, typename RowT>
__device__ __forceinline__ void engine_step(RowT (&row)[RPL], RowT (&disp)[RPL], int &sw, Piece &pc, int action,
                                            int lane, const Params &p, int e, int &reward, int &done,
                                            int &errbits, RowT walls)
{
    using M = typename CMask<RPL>::type;
    const int H = p.H;
    reward = p.reward_step;  // ref:256
    done = 0;
    if (pc.id >= 7) {  // no piece yet: the reference would fail on shape None (ref:170-172,245)
        errbits |= 4;
#pragma unroll
        for (int k = 0; k < RPL; ++k) disp[k] = row[k];
        return;
    }
    if (action > 6) { errbits |= 2; action = 6; }

    // ---- action (ref:245, 39-73): try the move, keep it unless it collides ----
    int r2 = pc.rot, x2 = pc.x;
    if (action == 0) x2 -= 1;
    if (action == 1) x2 += 1;
    if (action == 4) r2 = (r2 + 1) & 3;
    if (action == 5) r2 = (r2 + 3) & 3;
    PieceRows<RowT> pr = piece_rows<RowT>(pc.id, r2, x2);
    M cm = collision_mask<RPL, RowT>(row, pr, H);
    const bool moved = (r2 != pc.rot) || (x2 != pc.x);
    if (moved && ((cm >> pc.y) & 1)) {  // blocked: stay (ref:41,46,64,69)
        pr = piece_rows<RowT>(pc.id, pc.rot, pc.x);
        cm = collision_mask<RPL, RowT>(row, pr, H);
    } else {
        pc.rot = r2; pc.x = x2;
    }
    int y = pc.y;
    if (action == 3 && !((cm >> (y + 1)) & 1)) y += 1;  // soft_drop (ref:49-51)
    if (action == 2) {                                   // hard_drop (ref:54-59): first blocked height below
        const M above = cm >> (y + 1);
        if (above) y += ctz(above);
    }
    // ---- gravity (ref:247-250) ----
    int ld = get(sw, 1);
    if (!((cm >> (y + 1)) & 1)) {
        y += 1;
        if (p.step_reset) ld = 0;
    }
    pc.y = y;
    if (lane == 2) sw += 1;  // time += 1 (ref:253)

    // ---- grounded -> lock delay -> lock (ref:259-299) ----
    if ((cm >> (y + 1)) & 1) {
        ld += 1;  // ref:175,260: (x + 1) % (max(lock_delay, 0) + 1)
        if (ld >= p.lock_mod) ld %= p.lock_mod;
        if (ld == 0) engine_lock<RPL, RowT>(row, sw, pc, pr, lane, p, e, reward, done, errbits, walls);
    }
    put(sw, lane, 1, ld);
    // ---- compose the returned state (ref:301-303): draw, copy, erase ----
    RowT pm[RPL];
    piece_on_rows<RPL, RowT>(pm, pr, pc.y, lane);
#pragma unroll
    for (int k = 0; k < RPL; ++k) {
        disp[k] = row[k] | pm[k];
        row[k] = (row[k] & ~pm[k]) | walls;  // _set_piece(False) also wipes a just-locked piece after game over (ref:303)
    }
}

// ---------------------------------------------------------------------------------------------
// Observation writers
// ---------------------------------------------------------------------------------------------
// Four consecutive observation elements: one 16-byte store in the reference's float32, or one 4-byte store in
// the uint8 mode (same values: 0/1 for ram, 0/128/190 for images; reported separately, never as the parity mode).
__device__ __forceinline__ void store4(void *obs, size_t idx4, const float4 &v, bool u8)
{
    if (u8)
        reinterpret_cast<uchar4 *>(obs)[idx4] =
            make_uchar4((unsigned char)v.x, (unsigned char)v.y, (unsigned char)v.z, (unsigned char)v.w);
    else
        reinterpret_cast<float4 *>(obs)[idx4] = v;
}

// ram (ref:421-424 + float32 cast ref:400): out[x][y] = cell (x,y); rows come from the warp's smem row.
__device__ __forceinline__ void write_ram(const uint32_t *srow, void *out, const Params &p, int lane)
{
    const int H = p.H, W = p.W;
    const bool u8 = p.obs_u8 != 0;
    if ((H & 3) == 0) {
        const int nq = (W * H) >> 2, hq = H >> 2;
        for (int q = lane; q < nq; q += 32) {
            const int x = (int)(((uint32_t)q * p.inv_hq20) >> 20);
            const int yq = q - x * hq;
            const uint32_t bit = 1u << x;
            const uint4 r = reinterpret_cast<const uint4 *>(srow)[yq];
            store4(out, q, make_float4((r.x & bit) ? 1.0f : 0.0f, (r.y & bit) ? 1.0f : 0.0f, (r.z & bit) ? 1.0f : 0.0f,
                                       (r.w & bit) ? 1.0f : 0.0f), u8);
        }
    } else {
        const int nel = W * H;
        for (int i = lane; i < nel; i += 32) {
            const int x = (int)(((uint32_t)i * p.inv_h20) >> 20);
            const int yy = i - x * H;
            const bool on = ((srow[yy] >> x) & 1u) != 0u;
            if (u8) reinterpret_cast<unsigned char *>(out)[i] = on ? 1 : 0;
            else reinterpret_cast<float *>(out)[i] = on ? 1.0f : 0.0f;
        }
    }
}

// Per-thread constants of the image writers: this thread always writes float4 slot `k` of an image row.
struct ColSlot {
    float lo[4], hi[4];
    uint32_t mk[4];
};

template <int CH>
__device__ __forceinline__ ColSlot make_col_slot(int k, const Params &p)
{
    ColSlot cs;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = (4 * k + i) / CH;  // pixel column
        const int cc = c - p.pad_left;
        const bool inside = cc >= 0 && cc < p.inner_h;
        const bool cell = inside && (cc % p.pitch) >= p.gap;
        cs.lo[i] = inside ? 128.0f : 0.0f;            // background / border shade (ref:77-78)
        cs.hi[i] = cell ? 190.0f : cs.lo[i];           // piece shade (ref:79)
        cs.mk[i] = cell ? (1u << (cc / p.pitch)) : 0u;
    }
    return cs;
}

// ---------------------------------------------------------------------------------------------
// The step / reset / observe kernel.  OBS: 0 ram, 1 grayscale, 2 rgb.  MODE is a template parameter so
// that the step kernel carries no reset/observe code.
// ---------------------------------------------------------------------------------------------
// Warps (= envs) per CTA: image modes need 8 (252 writer threads); ram warps are independent, and smaller CTAs
// spread a small batch more evenly over the 148 SMs.
template <int OBS> struct Wpc { static constexpr int value = OBS == 0 ? kRamWarpsPerCta : kWarpsPerCta; };

template <int RPL, int OBS, int MODE, typename RowT, bool MANY>
__global__ void __launch_bounds__(32 * Wpc<OBS>::value, OBS == 0 ? ST_RAM_MINBLOCKS : ST_IMG_MINBLOCKS) st_main_kernel(const __grid_constant__ Params p)
{
    constexpr int WPC = Wpc<OBS>::value;
    __shared__ __align__(16) uint32_t s_disp[WPC][32 * RPL];
    __shared__ signed char s_rowy[OBS == 0 ? 1 : kImage];
    __shared__ unsigned char s_active[WPC];
    // terminal observation of envs that are auto-reset in this step (gym<=0.25 vector semantics: info["terminal_observation"])
    __shared__ __align__(16) uint32_t s_term[WPC][32 * RPL];
    __shared__ unsigned char s_has_term[WPC];
    // rgb: the CTA's output leaves through TMA bulk stores from a 3-deep ring of shared-memory chunks (measured
    // +5 % over direct 16-byte stores); grayscale and ram keep direct stores (bulk stores measured 8 % slower there)
    constexpr bool kBulk = (OBS == 2 || (OBS == 1 && ST_IMG_BULK_GRAY != 0)) && ST_IMG_BULK != 0;
    constexpr int kBulkPasses = ST_IMG_BULK_PASSES;
    __shared__ __align__(128) float4 s_bulk[kBulk ? 3 : 1][kBulk ? kBulkPasses * 252 : 1];

    // Programmatic dependent launch: let the next step's grid start its prologue while this one runs ...
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = (int)p.n;
    const int e = blockIdx.x * WPC + warp;  // this warp's env (launch_main keeps n below 2^31)
    const bool valid = e < n;
    if (OBS == 0 && !valid) return;  // ram warps never meet at a CTA barrier
    const RowT walls = (RowT)0xF | (~(RowT)0 << (p.W + OFF));
    constexpr int CH = OBS == 2 ? 3 : 1;
    constexpr int KPR = kImage * CH / 4;  // float4 slots per image row: 21 / 63
    constexpr int NG = 252 / KPR;         // row groups: 12 / 4
    ColSlot cs;
    int slot_k = 0, slot_g = 0;
    if (OBS != 0) {
        slot_k = threadIdx.x % KPR;
        slot_g = threadIdx.x / KPR;
        cs = make_col_slot<CH>(slot_k, p);
        if (threadIdx.x < kImage) {  // image row -> board row (-1 gap row, -2 border row)
            const int rr = (int)threadIdx.x - p.pad_top;
            s_rowy[threadIdx.x] = (rr < 0 || rr >= p.inner_v) ? -2 : ((rr % p.pitch) < p.gap ? -1 : rr / p.pitch);
        }
    }

    // ... and do not touch global memory before the previous grid in the stream has completed and flushed.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    bool selected = valid;
    // The action byte is the second cold miss of a step; issue its load first (asm volatile keeps it here) so that it
    // overlaps the state loads instead of queueing behind the first shuffle that consumes them.
    unsigned int action_u = 6u;
    if (MODE == MODE_STEP && valid) asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(action_u) : "l"(p.actions + e));
    if (MODE == MODE_RESET && valid && p.mask && p.mask[e] == 0) selected = false;
    unsigned char *rec = p.state + (size_t)(valid ? e : 0) * (unsigned)p.stride;
    int *rec_w = reinterpret_cast<int *>(rec);
    RowT row[RPL], disp[RPL], row_in[RPL];
    int sw = 0, errbits = 0;
    Piece pc = {7, 0, 0, 0};
    uint32_t clo = 0, chi = 0;  // lane x < W: column x of the env record (bit y = cell (x, y)); chi = rows 32..63
    if (selected) {
        if (lane < kStateWords) sw = rec_w[lane];
        const uint32_t *cols = reinterpret_cast<const uint32_t *>(rec_w + kStateWords);
        if (lane < p.W) {
            if (RPL == 1) clo = cols[lane];
            else { clo = cols[2 * lane]; chi = cols[2 * lane + 1]; }
        }
    }
    rows_from_columns<RPL, RowT>(row, clo, chi, p.W, lane, walls);
    // running output pointers (advance per step in st_step_many)
    const uint8_t *act_p = MODE == MODE_STEP ? p.actions + e : nullptr;
    float *rew_p = MODE == MODE_STEP ? p.reward + e : nullptr;
    uint8_t *done_p = MODE == MODE_STEP ? p.done + e : nullptr;
    int32_t *info_p = (MODE == MODE_STEP && p.info) ? p.info + (size_t)e * kStateWords + lane : nullptr;
    const bool u8 = p.obs_u8 != 0;
    const size_t esz = u8 ? 1 : 4;  // bytes per observation element
    char *obs_p = p.obs ? reinterpret_cast<char *>(p.obs) + (size_t)(OBS == 0 ? e : (int)blockIdx.x * WPC) * (unsigned)p.obs_elems * esz
                        : nullptr;
    char *term_p = (MODE == MODE_STEP && p.term_obs)
                       ? reinterpret_cast<char *>(p.term_obs) + (size_t)(OBS == 0 ? e : (int)blockIdx.x * WPC) * (unsigned)p.obs_elems * esz
                       : nullptr;
#if ST_ANCHOR_PTRS
    // Materialise the output addresses now, while the state loads are in flight (the compiler would otherwise sink
    // this arithmetic to the stores at the end of the step, behind the whole dependent chain).
    asm volatile("" : "+l"(rew_p), "+l"(done_p), "+l"(info_p), "+l"(obs_p));
#endif
    int action = (int)action_u;
    if (selected) pc = unpack_piece(get(sw, 0));
#pragma unroll
    for (int k = 0; k < RPL; ++k) { row_in[k] = row[k]; disp[k] = row[k]; }

    const int T = (MODE == MODE_STEP && MANY) ? p.T : 1;  // single-step launches compile without the step loop
    for (int t = 0; t < T; ++t) {
        bool has_term = false;
        if (selected) {
            if (MODE == MODE_STEP) {
                int reward = 0, done = 0;
                engine_step<RPL, RowT>(row, disp, sw, pc, action, lane, p, e, reward, done, errbits, walls);
                put(sw, lane, 0, pack_piece(pc));
                if (info_p && lane < kStateWords) *info_p = lane == 0 ? pc.id : sw;  // get_info (ref:232-241), pre-reset
                if (done) {
                    if (p.stats && lane >= 2 && lane <= 4)  // sum(time), sum(score), sum(lines) at done
                        atomicAdd(p.stats + (lane == 2 ? 1 : lane == 3 ? 3 : 2), (unsigned long long)(long long)sw);
                    if (p.stats && lane == 0) atomicAdd(p.stats, 1ull);
                    if (p.auto_reset) {  // VecEnv: reset obs = empty board, piece not drawn (ref:313-315)
                        if (term_p) {  // what step() returned in the reference at this terminal step (ref:301-302)
#pragma unroll
                            for (int k = 0; k < RPL; ++k) s_term[warp][lane + 32 * k] = (uint32_t)(disp[k] >> OFF) & p.fullmask;
                            has_term = true;
                        }
                        engine_clear<RPL, RowT>(row, sw, pc, lane, p, e, errbits, walls);
                        put(sw, lane, 0, pack_piece(pc));
#pragma unroll
                        for (int k = 0; k < RPL; ++k) disp[k] = walls;
                    }
                }
                if (lane == 0) {
                    *rew_p = (float)reward;
                    *done_p = (unsigned char)done;
                }
                if (MANY && t + 1 < T) {  // next step of st_step_many
                    act_p += n; rew_p += n; done_p += n;
                    if (info_p) info_p += p.info_t_stride;
                    action = *act_p;
                }
            } else if (MODE == MODE_RESET) {
                engine_clear<RPL, RowT>(row, sw, pc, lane, p, e, errbits, walls);
                put(sw, lane, 0, pack_piece(pc));
#pragma unroll
                for (int k = 0; k < RPL; ++k) disp[k] = walls;
            } else {  // MODE_OBSERVE: engine.render() (ref:317-321) or the bare board
                RowT pm[RPL];
#pragma unroll
                for (int k = 0; k < RPL; ++k) pm[k] = 0;
                if (p.draw_piece && pc.id < 7) {
                    const PieceRows<RowT> pr = piece_rows<RowT>(pc.id, pc.rot, pc.x);
                    piece_on_rows<RPL, RowT>(pm, pr, pc.y, lane);
                }
#pragma unroll
                for (int k = 0; k < RPL; ++k) disp[k] = row[k] | pm[k];
            }
            if (obs_p) {
#pragma unroll
                for (int k = 0; k < RPL; ++k) s_disp[warp][lane + 32 * k] = (uint32_t)(disp[k] >> OFF) & p.fullmask;
            }
        }
        if (OBS == 0) {
            __syncwarp();
            if (selected && obs_p) write_ram(s_disp[warp], obs_p, p, lane);  // unselected (masked-out) envs keep their obs
            if (has_term) write_ram(s_term[warp], term_p, p, lane);
            __syncwarp();
        } else {
            if (lane == 0) {
                s_active[warp] = selected && obs_p;
                s_has_term[warp] = has_term;
            }
            __syncthreads();
            // TMA bulk-store path: the CTA's 8 images are one contiguous region; it is produced in chunks of
            // kBulkPasses x 252 float4 in shared memory (3 buffers) and each chunk leaves with one
            // cp.async.bulk.global.shared::cta issued by thread 0.  Needs all 8 envs active (a chunk straddles
            // images); the tail CTA and masked resets take the direct-store path below.
            bool all_active = kBulk;
#pragma unroll
            for (int w = 0; w < WPC; ++w) all_active = all_active && s_active[w];
            if (kBulk && all_active) {
                constexpr int kPassF4 = KPR * NG;                         // 252 float4 per pass
                constexpr int kChunkF4 = kBulkPasses * kPassF4;
                constexpr int kChunks = (WPC * kImage / NG) / kBulkPasses;  // passes per CTA / passes per chunk
                const uint32_t chunk_bytes = (uint32_t)(kChunkF4 * (u8 ? sizeof(uchar4) : sizeof(float4)));
                for (int c = 0; c < kChunks; ++c) {
                    float4 *buf = s_bulk[c % 3];
                    if (threadIdx.x < kPassF4) {
#pragma unroll
                        for (int q = 0; q < kBulkPasses; ++q) {
                            const int rowg = slot_g + NG * (c * kBulkPasses + q);  // row index over the 8 images
                            const int w = rowg / kImage, rho = rowg - w * kImage;
                            const int code = s_rowy[rho];
                            const uint32_t b = code >= 0 ? s_disp[w][code] : 0u;
                            float4 v;
                            v.x = (b & cs.mk[0]) ? cs.hi[0] : cs.lo[0];
                            v.y = (b & cs.mk[1]) ? cs.hi[1] : cs.lo[1];
                            v.z = (b & cs.mk[2]) ? cs.hi[2] : cs.lo[2];
                            v.w = (b & cs.mk[3]) ? cs.hi[3] : cs.lo[3];
                            if (code == -2) v = make_float4(0.f, 0.f, 0.f, 0.f);
                            store4(buf, q * kPassF4 + threadIdx.x, v, u8);  // same ring, 4x denser in uint8 mode
                        }
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    }
                    __syncthreads();
                    if (threadIdx.x == 0) {
                        const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(buf);
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(obs_p + (size_t)c * chunk_bytes),
                                     "r"(saddr), "r"(chunk_bytes)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");  // chunk c-1 has left its buffer
                    }
                }
                if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncthreads();
            } else if (threadIdx.x < KPR * NG) {
                const size_t img4 = (size_t)(p.obs_elems >> 2);  // 4-element groups per image
#pragma unroll kImgUnroll
                for (int rho = slot_g; rho < kImage; rho += NG) {
                    const int code = s_rowy[rho];
#pragma unroll
                    for (int w = 0; w < WPC; ++w) {
                        if (!s_active[w]) continue;
                        const uint32_t b = code >= 0 ? s_disp[w][code] : 0u;
                        float4 v;
                        v.x = (b & cs.mk[0]) ? cs.hi[0] : cs.lo[0];
                        v.y = (b & cs.mk[1]) ? cs.hi[1] : cs.lo[1];
                        v.z = (b & cs.mk[2]) ? cs.hi[2] : cs.lo[2];
                        v.w = (b & cs.mk[3]) ? cs.hi[3] : cs.lo[3];
                        if (code == -2) v = make_float4(0.f, 0.f, 0.f, 0.f);
                        store4(obs_p, (size_t)w * img4 + rho * KPR + slot_k, v, u8);
                    }
                }
            }
            if (MODE == MODE_STEP && term_p && threadIdx.x < KPR * NG) {  // rare: images of the envs that just ended
                const size_t img4 = (size_t)(p.obs_elems >> 2);
#pragma unroll 1
                for (int w = 0; w < WPC; ++w) {
                    if (!s_has_term[w]) continue;
                    for (int rho = slot_g; rho < kImage; rho += NG) {
                        const int code = s_rowy[rho];
                        const uint32_t b = code >= 0 ? s_term[w][code] : 0u;
                        float4 v;
                        v.x = (b & cs.mk[0]) ? cs.hi[0] : cs.lo[0];
                        v.y = (b & cs.mk[1]) ? cs.hi[1] : cs.lo[1];
                        v.z = (b & cs.mk[2]) ? cs.hi[2] : cs.lo[2];
                        v.w = (b & cs.mk[3]) ? cs.hi[3] : cs.lo[3];
                        if (code == -2) v = make_float4(0.f, 0.f, 0.f, 0.f);
                        store4(term_p, (size_t)w * img4 + rho * KPR + slot_k, v, u8);
                    }
                }
            }
            __syncthreads();
        }
        if (MANY && term_p) term_p += p.obs_t_stride * (long long)esz;
        if (MANY && obs_p) obs_p += p.obs_t_stride * (long long)esz;
    }

    if (selected && MODE != MODE_OBSERVE) {
        if (lane < kStateWords) rec_w[lane] = sw;
        bool dirty = false;
#pragma unroll
        for (int k = 0; k < RPL; ++k) dirty |= row[k] != row_in[k];
        if (__any_sync(FULL, dirty)) {
            columns_from_rows<RPL, RowT>(row, clo, chi, p.W, lane);
            uint32_t *cols = reinterpret_cast<uint32_t *>(rec_w + kStateWords);
            if (lane < p.W) {
                if (RPL == 1) cols[lane] = clo;
                else { cols[2 * lane] = clo; cols[2 * lane + 1] = chi; }
            }
        }
    }
    if (errbits && p.err && lane == 0) atomicOr(p.err, errbits);
}

// ---------------------------------------------------------------------------------------------
// init / get_state / set_state: one thread per env, cold path.
// ---------------------------------------------------------------------------------------------
__global__ void st_init_kernel(const __grid_constant__ Params p)
{
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    int *w = reinterpret_cast<int *>(p.state + e * (long long)p.stride);
    for (int i = 0; i < p.stride / 4; ++i) w[i] = 0;
    w[0] = 7;    // no piece (ref:170-172)
    w[2] = -1;   // time  (ref:165)
    w[3] = -1;   // score (ref:166)
}

__global__ void st_get_state_kernel(const __grid_constant__ Params p, uint8_t *boards, int32_t *scalars)
{
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    const unsigned char *rec = p.state + e * (long long)p.stride;
    const int *w = reinterpret_cast<const int *>(rec);
    if (scalars) {
        int32_t *s = scalars + e * 18;
        const Piece pc = unpack_piece(w[0]);
        s[0] = pc.id; s[1] = pc.rot; s[2] = pc.x; s[3] = pc.y;
        for (int i = 1; i < kStateWords; ++i) s[3 + i] = w[i];
    }
    if (boards) {
        const uint32_t *cols = reinterpret_cast<const uint32_t *>(rec + 4 * kStateWords);
        for (int x = 0; x < p.W; ++x) {
            const unsigned long long c = p.col_words == 1 ? cols[x] : (cols[2 * x] | ((unsigned long long)cols[2 * x + 1] << 32));
            for (int y = 0; y < p.H; ++y) boards[(e * p.W + x) * p.H + y] = (uint8_t)((c >> y) & 1ull);
        }
    }
}

__global__ void st_set_state_kernel(const __grid_constant__ Params p, const uint8_t *boards, const int32_t *scalars)
{
    const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e >= p.n) return;
    unsigned char *rec = p.state + e * (long long)p.stride;
    int *w = reinterpret_cast<int *>(rec);
    if (scalars) {
        const int32_t *s = scalars + e * 18;
        Piece pc;
        pc.id = s[0] < 0 || s[0] > 7 ? 7 : s[0];
        pc.rot = s[1] & 3;
        pc.x = min(max(s[2], 0), p.W - 1);
        pc.y = min(max(s[3], 0), p.H - 1);
        w[0] = pack_piece(pc);
        for (int i = 1; i < kStateWords; ++i) w[i] = s[3 + i];
    }
    if (boards) {
        uint32_t *cols = reinterpret_cast<uint32_t *>(rec + 4 * kStateWords);
        for (int x = 0; x < p.W; ++x) {
            unsigned long long c = 0;
            for (int y = 0; y < p.H; ++y) c |= (boards[(e * p.W + x) * p.H + y] ? 1ull : 0ull) << y;
            if (p.col_words == 1) cols[x] = (uint32_t)c;
            else { cols[2 * x] = (uint32_t)c; cols[2 * x + 1] = (uint32_t)(c >> 32); }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// K3: TetrisEnv.render('rgb_array') (ref:458-462): engine.render() -> convert_grayscale(obs, size) ->
// convert_grayscale_rgb, uint8 [n][size][size][3].  One CTA per env; display-only path, kept simple.
// ---------------------------------------------------------------------------------------------
struct RenderGeom {
    int size, pitch, gap, inner_v, inner_h, pad_top, pad_left;
};

__device__ __forceinline__ uint32_t shade_at(const uint32_t *disp, const RenderGeom &g, int pix)
{
    const int r = pix / g.size, c = pix - r * g.size;
    const int rr = r - g.pad_top, cc = c - g.pad_left;
    if (rr < 0 || rr >= g.inner_v || cc < 0 || cc >= g.inner_h) return 0u;   // border_shade (ref:77)
    if ((rr % g.pitch) < g.gap || (cc % g.pitch) < g.gap) return 128u;       // background_shade (ref:78)
    return ((disp[rr / g.pitch] >> (cc / g.pitch)) & 1u) ? 190u : 128u;      // piece_shade (ref:79)
}

__global__ void __launch_bounds__(256) st_render_kernel(const __grid_constant__ Params p, const RenderGeom g, uint8_t *out)
{
    __shared__ uint32_t disp[64];
    const long long e = blockIdx.x;
    const unsigned char *rec = p.state + e * (long long)p.stride;
    if (threadIdx.x < 64) {
        const int Y = threadIdx.x;
        uint32_t v = 0;
        if (Y < p.H) {
            const uint32_t *cols = reinterpret_cast<const uint32_t *>(rec + 4 * kStateWords);
            for (int x = 0; x < p.W; ++x) {
                const uint32_t c = p.col_words == 1 ? cols[x] : cols[2 * x + (Y >> 5)];
                v |= ((c >> (Y & 31)) & 1u) << x;
            }
            const Piece pc = unpack_piece(*reinterpret_cast<const int *>(rec));
            if (p.draw_piece && pc.id < 7) {  // _set_piece(True) (ref:323-327): in-board cells only
                const PieceRows<unsigned long long> pr = piece_rows<unsigned long long>(pc.id, pc.rot, pc.x);
                const int t = Y - pc.y - pr.minj;
                if (t >= 0 && t < 4) v |= (uint32_t)(pr.m[t] >> OFF) & p.fullmask;
            }
        }
        disp[Y] = v;
    }
    __syncthreads();
    const long long nbytes = (long long)g.size * g.size * 3;
    uint8_t *o = out + e * nbytes;
    if ((nbytes & 3) == 0) {
        uint32_t *o4 = reinterpret_cast<uint32_t *>(o);
        for (int w = threadIdx.x; w < (int)(nbytes >> 2); w += blockDim.x) {
            const int b = 4 * w;
            o4[w] = shade_at(disp, g, b / 3) | (shade_at(disp, g, (b + 1) / 3) << 8) |
                    (shade_at(disp, g, (b + 2) / 3) << 16) | (shade_at(disp, g, (b + 3) / 3) << 24);
        }
    } else {
        for (int b = threadIdx.x; b < (int)nbytes; b += blockDim.x) o[b] = (uint8_t)shade_at(disp, g, b / 3);
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
#ifndef ST_SMALL_RAM_BATCH
#define ST_SMALL_RAM_BATCH 6144
#endif
constexpr long long kSmallRamBatch = ST_SMALL_RAM_BATCH;
static unsigned long long g_launches = 0;
unsigned long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }
static inline void count_launch() { __atomic_add_fetch(&g_launches, 1ull, __ATOMIC_RELAXED); }

template <int RPL, int OBS, int MODE, typename RowT, bool MANY>
static cudaError_t launch_t(const Params &p, cudaStream_t stream)
{
    constexpr int WPC = Wpc<OBS>::value;
    const long long ngroups = (p.n + WPC - 1) / WPC;
    if (ngroups == 0) return cudaSuccess;
    static const bool pdl = getenv("ST_B200_NO_PDL") == nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ngroups);
    cfg.blockDim = dim3(32 * WPC);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    if (OBS != 0) {
        // Unused dynamic shared memory caps the resident CTAs per SM.  Measured on B200: the direct-store grayscale
        // writer is fastest with 2 CTAs per SM (1.044 ms at 262144 envs; 1.094 with 3, 1.143 with 4: fewer concurrent
        // write streams keep DRAM pages open longer), the bulk-store rgb writer with 4 (its own 24 KB ring, no pad).
        static const int dsmem = getenv("ST_B200_IMG_DSMEM") ? atoi(getenv("ST_B200_IMG_DSMEM")) : (OBS == 1 ? 100000 : 0);
        if (dsmem > 0) {
            static thread_local int attr_set_for_device = -1;  // per instantiation: opt in to > 48 KB once per device
            int dev = 0;
            cudaGetDevice(&dev);
            if (attr_set_for_device != dev) {
                cudaFuncSetAttribute(st_main_kernel<RPL, OBS, MODE, RowT, MANY>, cudaFuncAttributeMaxDynamicSharedMemorySize, dsmem);
                attr_set_for_device = dev;
            }
            cfg.dynamicSmemBytes = (size_t)dsmem;
        }
    }
    count_launch();
    return cudaLaunchKernelEx(&cfg, st_main_kernel<RPL, OBS, MODE, RowT, MANY>, p);
}

}  // namespace st
#include "st_kernels_tpe.cuh"
#include "st_kernels_cols.cuh"
namespace st {

template <int RPL, int OBS, typename RowT>
static cudaError_t launch_mode(const Params &p, cudaStream_t stream)
{
    switch (p.mode) {
    case MODE_STEP:
        // The looped build keeps more values live (72 vs 47 registers) and overlaps more address arithmetic with the
        // state loads: 0.4-1 us faster per launch on latency-bound ram batches (<= ~6k envs), while the loop-free
        // build executes ~25 % fewer instructions and wins on everything larger (measured, tools/warp_variant_sweep.py).
        return (p.T > 1 || ST_FORCE_MANY || (OBS == 0 && p.n <= kSmallRamBatch))
                   ? launch_t<RPL, OBS, MODE_STEP, RowT, true>(p, stream)
                   : launch_t<RPL, OBS, MODE_STEP, RowT, false>(p, stream);
    case MODE_RESET: return launch_t<RPL, OBS, MODE_RESET, RowT, false>(p, stream);
    case MODE_OBSERVE: return launch_t<RPL, OBS, MODE_OBSERVE, RowT, false>(p, stream);
    }
    return cudaErrorInvalidValue;
}

template <int OBS>
static cudaError_t launch_obs(const Params &p, cudaStream_t stream)
{
    const bool two = p.H > 31, wide = p.W + OFF + 3 > 31;  // piece bits reach column W + 2
    if (!two && !wide) return launch_mode<1, OBS, uint32_t>(p, stream);
    if (two && !wide) return launch_mode<2, OBS, uint32_t>(p, stream);
    if (!two && wide) return launch_mode<1, OBS, unsigned long long>(p, stream);
    return launch_mode<2, OBS, unsigned long long>(p, stream);
}

// Which kernel steps a ram batch: 0 = warp-per-env on row lanes (K1, also every image mode / reset / observe launch),
// 1 = thread-per-env (K1b), 2 = warp-per-env on column lanes (K1c).  ST_B200_RAM_PATH = warp | thread | cols forces one
// (where eligible); auto: thread-per-env from tpe_min_envs() envs up, where a warp per env is issue-bound, column lanes
// below (a warp per env has the shorter critical path there), row lanes for boards wider than 24 columns.
static int ram_path(const Params &p, int obs_type)
{
    const char *path = getenv("ST_B200_RAM_PATH");
    const char f = path ? path[0] : 'a';
    const bool tpe_ok = tpe_eligible(p, obs_type), cols_ok = cols_eligible(p, obs_type);
    if (f == 't') return tpe_ok ? 1 : cols_ok ? 2 : 0;
    if (f == 'w') return 0;
    if (f == 'c') return cols_ok ? 2 : 0;
    if (tpe_ok && p.n >= tpe_min_envs(p)) return 1;
    return cols_ok ? 2 : 0;
}

cudaError_t launch_main(const Params &p, int obs_type, cudaStream_t stream)
{
    if (p.n >= (1ll << 31) - 8) return cudaErrorInvalidValue;
    switch (ram_path(p, obs_type)) {
    case 1: return launch_tpe(p, stream);
    case 2: return launch_cols(p, stream);
    }
    switch (obs_type) {
    case 0: return launch_obs<0>(p, stream);
    case 1: return launch_obs<1>(p, stream);
    case 2: return launch_obs<2>(p, stream);
    }
    return cudaErrorInvalidValue;
}

// Which kernel a single-step launch of this configuration runs (for reports and profiles).
const char *step_kernel_name(const Params &p, int obs_type)
{
    Params q = p;
    q.mode = MODE_STEP;
    q.T = 1;
    switch (ram_path(q, obs_type)) {
    case 1: return "st_step_tpe_kernel";
    case 2: return "st_step_cols_kernel";
    }
    return obs_type == 0 ? "st_main_kernel<ram,STEP>" : obs_type == 1 ? "st_main_kernel<grayscale,STEP>" : "st_main_kernel<rgb,STEP>";
}

cudaError_t launch_render(const Params &p, int size, uint8_t *out, cudaStream_t stream)
{
    if (p.n == 0) return cudaSuccess;
    RenderGeom g;
    const int limiting = p.W > p.H ? p.W : p.H;
    g.size = size;
    g.gap = size / 100 + 1;                                   // ref:87
    const int bs = (size - 2 * g.gap) / limiting - g.gap;     // ref:88
    if (bs < 0) return cudaErrorInvalidValue;                 // the reference raises in np.repeat
    g.pitch = bs + g.gap;
    g.inner_v = g.gap + g.pitch * p.H;                        // ref:90-91 on the transposed (H, W) array
    g.inner_h = g.gap + g.pitch * p.W;
    g.pad_top = (size - g.inner_v) / 2;                       // ref:93-94
    g.pad_left = (size - g.inner_h) / 2;
    st_render_kernel<<<(unsigned)p.n, 256, 0, stream>>>(p, g, out);
    count_launch();
    return cudaGetLastError();
}

static inline unsigned cold_grid(long long n) { return (unsigned)((n + 127) / 128); }

cudaError_t launch_init(const Params &p, cudaStream_t stream)
{
    if (p.n == 0) return cudaSuccess;
    st_init_kernel<<<cold_grid(p.n), 128, 0, stream>>>(p);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_get_state(const Params &p, uint8_t *boards, int32_t *scalars, cudaStream_t stream)
{
    if (p.n == 0) return cudaSuccess;
    st_get_state_kernel<<<cold_grid(p.n), 128, 0, stream>>>(p, boards, scalars);
    count_launch();
    return cudaGetLastError();
}
cudaError_t launch_set_state(const Params &p, const uint8_t *boards, const int32_t *scalars, cudaStream_t stream)
{
    if (p.n == 0) return cudaSuccess;
    st_set_state_kernel<<<cold_grid(p.n), 128, 0, stream>>>(p, boards, scalars);
    count_launch();
    return cudaGetLastError();
}

}  // namespace st
