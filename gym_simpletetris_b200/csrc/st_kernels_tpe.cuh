// K1b — thread-per-env ram step for LARGE batches (included by st_kernels.cu).
//
// The warp-per-env kernel (K1) spends ~400 warp-instructions per env-step: ideal for small batches, where the
// machine is latency-bound and a whole warp per env keeps every rare branch uniform, but issue-bound from
// ~16k envs up.  Here one THREAD steps one env and the warp does the memory work cooperatively:
//   1. the warp copies its 32 env records (32 x stride bytes, contiguous) from HBM into shared memory with
//      coalesced 4-byte loads; record pitch in smem is odd, so lane r touching word c of ITS record is
//      conflict-free;
//   2. each lane runs TetrisEngine.step (ref:243-304) on its record in smem — same rules, same widened-row
//      masks and Philox stream as K1, expressed per thread: a 7-row register window answers the collision
//      test for anchors y..y+3 (action, soft drop, gravity, grounded), hard drop slides a 4-row window down;
//   3. info / reward / done go out coalesced, auto-reset envs are cleared, every other lane ORs its piece
//      into its smem rows (the reference's _set_piece(True), ref:301);
//   4. the warp expands the 32 boards into float32 [W][H] with 16-byte stores, 512 contiguous bytes per warp
//      instruction; 5. lanes erase their piece again (ref:303) and the records are stored back, coalesced.
// About 55 warp-instructions per env-step, so ram mode becomes HBM-bound instead of issue-bound.
#pragma once

namespace st {

#ifndef ST_TPE_WARPS
#define ST_TPE_WARPS 4
#endif
constexpr int kTpeWarps = ST_TPE_WARPS;

template <typename RowT, bool ROWS16>
struct TpeRec {
    uint32_t *w;  // word 0 of this env's record in shared memory
    __device__ __forceinline__ uint32_t raw(int Y) const
    {
        return ROWS16 ? (uint32_t)reinterpret_cast<const uint16_t *>(w + kStateWords)[Y] : w[kStateWords + Y];
    }
    __device__ __forceinline__ void set_raw(int Y, uint32_t v) const
    {
        if (ROWS16) reinterpret_cast<uint16_t *>(w + kStateWords)[Y] = (uint16_t)v;
        else w[kStateWords + Y] = v;
    }
    // widened row as the collision test sees it: rows above the board are exempt from board AND walls
    // (ref:32-33) -> 0; rows below the floor collide with anything (ref:34) -> all ones
    __device__ __forceinline__ RowT widened(int Y, int H, RowT walls) const
    {
        const int Yc = Y < 0 ? 0 : (Y >= H ? H - 1 : Y);
        const RowT v = ((RowT)raw(Yc) << OFF) | walls;
        return Y < 0 ? (RowT)0 : (Y >= H ? ~(RowT)0 : v);
    }
};

template <typename RowT>
__device__ __forceinline__ PieceRows<RowT> tpe_piece_rows(const unsigned long long *s_tab, int id, int rot, int x)
{
    const unsigned long long e = s_tab[id * 4 + rot];  // smem copy: lanes index different entries
    const uint32_t lo = (uint32_t)e, hi = (uint32_t)(e >> 32);
    PieceRows<RowT> pr;
    pr.minj = (int)(hi & 15u) - 3;
    pr.maxj = (int)((hi >> 4) & 15u) - 3;
#pragma unroll
    for (int t = 0; t < 4; ++t) pr.m[t] = (RowT)((lo >> (7 * t)) & 127u) << (x + 1);
    return pr;
}

// bit d (0..3) set <=> is_occupied(shape, (x, y + d), board) (ref:29-36)
template <typename RowT, bool ROWS16>
__device__ __forceinline__ uint32_t tpe_collisions(const TpeRec<RowT, ROWS16> &rec, const PieceRows<RowT> &pr, int y,
                                                   int H, RowT walls)
{
    RowT R[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) R[i] = rec.widened(y + pr.minj + i, H, walls);
    uint32_t cm = 0;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const RowT hit = (pr.m[0] & R[d]) | (pr.m[1] & R[d + 1]) | (pr.m[2] & R[d + 2]) | (pr.m[3] & R[d + 3]);
        cm |= (hit != 0 ? 1u : 0u) << d;
    }
    return cm;
}

// _new_piece / _choose_shape (ref:183-200) on the record's shape_counts (words 8..14).
template <typename RowT, bool ROWS16>
__device__ __forceinline__ int tpe_spawn(const TpeRec<RowT, ROWS16> &rec, const Params &p, int e, int &errbits)
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
template <typename RowT, bool ROWS16>
__device__ __forceinline__ void tpe_lock(const TpeRec<RowT, ROWS16> &rec, Piece &pc, const PieceRows<RowT> &pr,
                                         const Params &p, int e, int &reward, int &done, int &errbits)
{
    const int H = p.H;
    const uint32_t fullmask = p.fullmask;
#pragma unroll
    for (int t = 0; t < 4; ++t) {  // _set_piece(True) (ref:263): in-board cells only
        const int Y = pc.y + pr.minj + t;
        if (Y >= 0 && Y < H) rec.set_raw(Y, rec.raw(Y) | ((uint32_t)(pr.m[t] >> OFF) & fullmask));
    }
    // One pass over the rows answers _clear_lines' can_clear (ref:206), _count_holes (ref:218-220) and
    // sum(np.any(board, axis=0)) (ref:287,289); only when a row really is full (rare) the board is compacted
    // (ref:207-214) and the holes / height pass repeated on the new board.
    int k = 0, holes = 0, nonempty = 0;
    uint32_t above = 0;
    for (int i = 0; i < H; ++i) {
        const uint32_t r = rec.raw(i);
        k += r == fullmask ? 1 : 0;
        holes += __popc(above & ~r & fullmask);
        above |= r;
        nonempty += r != 0u ? 1 : 0;
    }
    if (k) {
        int j = H - 1;
        for (int i = H - 1; i >= 0; --i) {
            const uint32_t r = rec.raw(i);
            if (r != fullmask) { rec.set_raw(j, r); --j; }
        }
        for (; j >= 0; --j) rec.set_raw(j, 0u);
        rec.w[4] += (uint32_t)k;
        holes = 0; nonempty = 0; above = 0;
        for (int i = 0; i < H; ++i) {
            const uint32_t r = rec.raw(i);
            holes += __popc(above & ~r & fullmask);
            above |= r;
            nonempty += r != 0u ? 1 : 0;
        }
    }
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
    if (rec.raw(0) != 0u) {  // ref:277-281
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
        pc.rot = 0; pc.x = p.W / 2; pc.y = 0;
    }
}

// TetrisEngine.step (ref:243-304) up to, not including, the composition of the returned state.
template <typename RowT, bool ROWS16>
__device__ __forceinline__ void tpe_engine_step(const TpeRec<RowT, ROWS16> &rec, int action, const Params &p, int e,
                                                RowT walls, const unsigned long long *s_tab, int &reward, int &done,
                                                int &errbits)
{
    const int H = p.H;
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
    PieceRows<RowT> pr = tpe_piece_rows<RowT>(s_tab, pc.id, r2, x2);
    uint32_t cm = tpe_collisions(rec, pr, pc.y, H, walls);
    const bool moved = (r2 != pc.rot) || (x2 != pc.x);
    if (moved && (cm & 1u)) {  // blocked: stay (ref:41,46,64,69)
        pr = tpe_piece_rows<RowT>(s_tab, pc.id, pc.rot, pc.x);
        cm = tpe_collisions(rec, pr, pc.y, H, walls);
    } else {
        pc.rot = r2; pc.x = x2;
    }
    int y = pc.y;
    int ld = (int)rec.w[1];
    bool grounded;
#ifdef ST_TPE_WHATIF_NODROP  // timing experiments only
    if (action == 2) action = 6;
#endif
    if (action == 2) {  // hard_drop (ref:54-59): first colliding anchor below, four candidates per window
        uint32_t c4 = cm >> 1;  // anchors y+1 .. y+3 are already known
        int ya = y + 1;
        while (c4 == 0u) {      // ends at the floor at the latest (rows >= H collide with everything)
            ya += (ya == y + 1) ? 3 : 4;
            c4 = tpe_collisions(rec, pr, ya, H, walls);
        }
        y = ya + (__ffs((int)c4) - 1) - 1;
        grounded = true;  // gravity (ref:247) cannot move it further, so step_reset does not fire
    } else {
        int d = 0;
        if (action == 3 && !((cm >> 1) & 1u)) d = 1;  // soft_drop (ref:49-51)
        if (!((cm >> (d + 1)) & 1u)) {                 // gravity (ref:247-250)
            d += 1;
            if (p.step_reset) ld = 0;
        }
        grounded = ((cm >> (d + 1)) & 1u) != 0u;       // _has_dropped (ref:202-203)
        y += d;
    }
    pc.y = y;
    rec.w[2] += 1u;  // time (ref:253)
    if (grounded) {
        ld += 1;
        if (ld >= p.lock_mod) ld %= p.lock_mod;
#ifndef ST_TPE_WHATIF_NOLOCK  // timing experiments only
        if (ld == 0) tpe_lock(rec, pc, pr, p, e, reward, done, errbits);
#endif
    }
    rec.w[1] = (uint32_t)ld;
    rec.w[0] = (uint32_t)pack_piece(pc);
}

template <typename RowT, bool ROWS16>
__global__ void __launch_bounds__(32 * kTpeWarps) st_step_tpe_kernel(const __grid_constant__ Params p)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];
    __shared__ unsigned long long s_tab[28];
    asm volatile("griddepcontrol.launch_dependents;");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < 28) s_tab[threadIdx.x] = c_tab.e[threadIdx.x];
    const int H = p.H, W = p.W;
    const int SW = p.stride >> 2;  // words per record in HBM
    const int pitch = SW | 1;      // odd pitch in smem: lane r, word c -> bank (r * pitch + c) % 32, conflict-free
    const int epw = p.tpe_epw;     // envs per warp (32, 16 or 8): fewer envs per warp = more warps for mid-size batches
    uint32_t *recs = s_dyn + warp * epw * pitch;
    const RowT walls = (RowT)0xF | (~(RowT)0 << (W + OFF));
    const long long e0 = ((long long)blockIdx.x * kTpeWarps + warp) * epw;
    int nvalid = (int)(p.n - e0 < epw ? p.n - e0 : epw);
    nvalid = nvalid < 0 ? 0 : nvalid;
    __syncthreads();  // s_tab
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (nvalid == 0) return;

    // the action bytes are issued first so that their miss overlaps the record copy
    const int e = (int)e0 + lane;
    unsigned int action_u = 6u;
    if (lane < nvalid) asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(action_u) : "l"(p.actions + e));

    // 1. records HBM -> smem
    uint32_t *g_rec = reinterpret_cast<uint32_t *>(p.state + e0 * (long long)p.stride);
    const int nwords = nvalid * SW;
    if (pitch == SW) {  // contiguous in both: 16-byte copies (32 records always start 128-byte aligned)
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

    const TpeRec<RowT, ROWS16> rec = {recs + lane * pitch};
    int errbits = 0;
    const size_t n_envs = (size_t)p.n;
    for (int t = 0; t < p.T; ++t) {  // st_step_many: the records stay in shared memory between steps
    // 2. engine, one env per lane
    int reward = 0, done = 0;
    if (lane < nvalid) tpe_engine_step(rec, (int)action_u, p, e, walls, s_tab, reward, done, errbits);
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
    if (p.term_obs) {  // terminal observation of the envs that end here: their board already holds the locked piece
        unsigned term = __ballot_sync(FULL, lane < nvalid && done && p.auto_reset);
        const bool u8 = p.obs_u8 != 0;
        char *tb = reinterpret_cast<char *>(p.term_obs) + ((long long)t * p.obs_t_stride + e0 * (long long)p.obs_elems) * (u8 ? 1 : 4);
        const int nel = W * H;
        while (term) {
            const int r = __ffs((int)term) - 1;
            term &= term - 1;
            const TpeRec<RowT, ROWS16> rr = {recs + r * pitch};
            char *dst = tb + (size_t)r * nel * (u8 ? 1 : 4);
            for (int i = lane; i < nel; i += 32) {
                const int x = (int)(((uint32_t)i * p.inv_h20) >> 20);
                const int yy = i - x * H;
                const bool on = ((rr.raw(yy) >> x) & 1u) != 0u;
                if (u8) reinterpret_cast<unsigned char *>(dst)[i] = on ? 1 : 0;
                else reinterpret_cast<float *>(dst)[i] = on ? 1.0f : 0.0f;
            }
        }
        __syncwarp();
    }
    uint32_t pbits[4] = {0u, 0u, 0u, 0u};
    int ptop = 0;
    if (lane < nvalid) {
        if (done && p.auto_reset) {  // clear() (ref:306-315): the reset observation is the empty board
#pragma unroll
            for (int i = 2; i <= 6; ++i) rec.w[i] = 0u;
            Piece pc;
            pc.id = tpe_spawn(rec, p, e, errbits);
            pc.rot = 0; pc.x = W / 2; pc.y = 0;
            rec.w[0] = (uint32_t)pack_piece(pc);
            for (int i = 0; i < H; ++i) rec.set_raw(i, 0u);
        } else {
            const Piece pc = unpack_piece((int)rec.w[0]);
            if (pc.id < 7) {  // _set_piece(True) (ref:301)
                const PieceRows<RowT> pr = tpe_piece_rows<RowT>(s_tab, pc.id, pc.rot, pc.x);
                ptop = pc.y + pr.minj;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int Y = ptop + t;
                    pbits[t] = (Y >= 0 && Y < H) ? ((uint32_t)(pr.m[t] >> OFF) & p.fullmask) : 0u;
                    if (pbits[t]) rec.set_raw(Y, rec.raw(Y) | pbits[t]);
                }
            }
        }
    }
    __syncwarp();

    // 4. observations: float32 [W][H] per env (ref:421-424, 400); 16-byte stores when H % 4 == 0, else 4-byte ones
    if (p.obs && (H & 3) != 0) {
        const int nel = W * H;
        const bool u8 = p.obs_u8 != 0;
        char *o = reinterpret_cast<char *>(p.obs) + ((long long)t * p.obs_t_stride + e0 * (long long)p.obs_elems) * (u8 ? 1 : 4);
        for (int i0 = 0; i0 < nel; i0 += 32) {
            const int i = i0 + lane;
            const int x = (int)(((uint32_t)i * p.inv_h20) >> 20);
            const int yy = i - x * H;
            if (i < nel) {
                for (int r = 0; r < nvalid; ++r) {
                    const TpeRec<RowT, ROWS16> rr = {recs + r * pitch};
                    const bool on = ((rr.raw(yy) >> x) & 1u) != 0u;
                    if (u8) reinterpret_cast<unsigned char *>(o)[r * nel + i] = on ? 1 : 0;
                    else reinterpret_cast<float *>(o)[r * nel + i] = on ? 1.0f : 0.0f;
                }
            }
        }
    } else if (p.obs) {
        const int hq = H >> 2, nq = W * hq;
        const bool u8 = p.obs_u8 != 0;
        char *o4 = reinterpret_cast<char *>(p.obs) + ((long long)t * p.obs_t_stride + e0 * (long long)p.obs_elems) * (u8 ? 1 : 4);
        for (int q0 = 0; q0 < nq; q0 += 32) {
            const int q = q0 + lane;
            const int x = (int)(((uint32_t)q * p.inv_hq20) >> 20);
            const int yq = q - x * hq;
            const uint32_t bit = 1u << x;
            if (q < nq) {
                for (int r = 0; r < nvalid; ++r) {
                    const uint32_t *rw = recs + r * pitch + kStateWords;
                    float4 v;
                    if (ROWS16) {  // two 16-bit rows per word: test bit x and bit x + 16 in place
                        const uint32_t w01 = rw[2 * yq], w23 = rw[2 * yq + 1], bith = bit << 16;
                        v = make_float4((w01 & bit) ? 1.0f : 0.0f, (w01 & bith) ? 1.0f : 0.0f,
                                        (w23 & bit) ? 1.0f : 0.0f, (w23 & bith) ? 1.0f : 0.0f);
                    } else {
                        v = make_float4((rw[4 * yq] & bit) ? 1.0f : 0.0f, (rw[4 * yq + 1] & bit) ? 1.0f : 0.0f,
                                        (rw[4 * yq + 2] & bit) ? 1.0f : 0.0f, (rw[4 * yq + 3] & bit) ? 1.0f : 0.0f);
                    }
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
            if (pbits[k]) rec.set_raw(ptop + k, rec.raw(ptop + k) & ~pbits[k]);
    }
    __syncwarp();
    }  // for t
    if (pitch == SW) {
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

template <typename RowT, bool ROWS16>
static cudaError_t launch_tpe_t(const Params &p, cudaStream_t stream)
{
    const long long nwarps = (p.n + p.tpe_epw - 1) / p.tpe_epw;
    const long long nctas = (nwarps + kTpeWarps - 1) / kTpeWarps;
    const int pitch = (p.stride >> 2) | 1;
    const size_t smem = (size_t)kTpeWarps * p.tpe_epw * pitch * 4;
    static const bool pdl = getenv("ST_B200_NO_PDL") == nullptr;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(st_step_tpe_kernel<RowT, ROWS16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
    return cudaLaunchKernelEx(&cfg, st_step_tpe_kernel<RowT, ROWS16>, p);
}

// Measured on B200 (tools/ram_path_sweep.py, us per step; 10x20 / 20 wide x 40 high boards):
//   n        warp    8/warp  16/warp  32/warp   |   warp    8/warp  16/warp  32/warp
//   8192     8.2     12.5    11.9     11.2      |   10.7    22.3    20.9     21.3
//   16384    12.0    15.0    15.6     15.3      |   16.6    27.9    35.7     37.5
//   32768    19.6    16.2    19.5     21.7      |   28.0    31.9    46.2     56.0
//   65536    34.8    22.6    20.5     27.7      |   50.8    49.1    59.2     73.4
//   262144   125.5   69.3    55.6     54.0      |  187.4   169.8   181.5    197.0
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
    const bool wide = p.W + OFF + 3 > 31;
    if (p.row_bytes == 2) return launch_tpe_t<uint32_t, true>(p, stream);  // W <= 16 is never wide
    if (!wide) return launch_tpe_t<uint32_t, false>(p, stream);
    return launch_tpe_t<unsigned long long, false>(p, stream);
}

}  // namespace st
