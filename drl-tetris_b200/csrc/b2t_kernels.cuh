// b200tetris — sm_100a kernels.  One thread = one game (reset / step / observe) or one simulated placement (expand);
// boards in shared memory (thread-private columns, stride = block size, so every access is bank-conflict free),
// scalars in registers, HBM traffic = coalesced 128-bit state vectors in, state + uint8 observation tensors out.
// The observation bytes are produced by a warp-cooperative pass: the 32 rows' bit streams are packed back to back
// in shared memory and every lane expands 16 bits -> one 128-bit store, so the [rows,2,H,W] tensor is written fully
// coalesced no matter how H*W aligns.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include "b2t_game.cuh"
#include "b2t_raw.cuh"
#include "b2t_rng.cuh"

namespace b2t {

#ifndef B2T_TPB
#define B2T_TPB 128
#endif
#ifndef B2T_GAME_MINB
#define B2T_GAME_MINB (512 / B2T_TPB)   // resident blocks per SM the step/expand kernel is register-capped for (16 warps)
#endif
constexpr int TPB = B2T_TPB;
constexpr int WARPS = TPB / 32;
constexpr int STAT_SLICES = 64;
constexpr int STAT_WORDS = 8;

struct DevTabs {
    const uint8_t* piece;
    const uint8_t* hole;
    const double* powtab;
    uint32_t piece_len, hole_len, spawn_rots;
    uint16_t pcn[28], pcc[28];
};

enum : uint32_t {
    KF_RESET = 1u,      // reset listed games with p.seeds before anything else
    KF_STEP = 2u,       // perform_action
    KF_FINALIZE = 4u,   // ... including finish_action
    KF_STORE = 8u,      // write the state back
    KF_EXPAND = 16u,    // rows are placements (offsets/placement cache), never stores
    KF_POKE_DEAD = 32u, // state.lock()-style write: p.poke_player / p.poke_value
    KF_CLEAR_FIELD = 64u,
    KF_INIT = 128u      // gamen_kernel: a never-reset game instead of the stored one (b2t_create)
};

struct KParams {
    uint4* state;
    uint32_t* tiles;              // optional colour planes [n_tiles][2*TILE_PLANES*SW][32] (B2T_FLAG_TILES), generic variant only
    int32_t n_games;
    const int32_t* idx;
    int32_t game0;                // idx == NULL: row i is game game0 + i (a slice of a pipelined host call)
    int32_t n;
    uint32_t flags;
    const uint32_t* seeds;        // KF_RESET
    const uint8_t* player;        // KF_STEP / KF_EXPAND
    const uint8_t* keys;
    const uint8_t* key_len;       // NULL: rows are zero-terminated (the length is the position of the last non-zero key)
    int32_t key_fmt;              // 0: one key code per byte; 1: 4-bit tokens, two per byte (low nibble first), lmax/2 bytes per row
    int32_t auto_reset;           // fused reset with engine-chosen seeds (auto_seed) when reset_seeds == NULL
    uint32_t seed_base;
    int32_t lmax;
    const int32_t* placement;
    const uint32_t* plc;          // placement cache [n_games][KMAX][KEYW]
    const int32_t* kcache;        // [n_games]
    int32_t elapsed;
    const uint32_t* reset_seeds;  // fused reset of finished games
    const int32_t* offsets;       // KF_EXPAND: [n+1]
    const int4* row_meta;         // KF_EXPAND: [capacity] {board slot i, game id, placement k, player} of every row
    int32_t capacity;
    int32_t poke_player, poke_value;
    // outputs (row-indexed)
    int8_t* reward;
    uint8_t* done;
    uint8_t* lines;
    uint8_t* field;
    uint8_t* vec;
    uint8_t* piece_idx;
    int32_t* scalars;
    int32_t obs_dtype;            // b2t_obs.dtype of field / vec: 0 uint8, 1 half, 2 bfloat16, 3 float, 4 fields as bits + uint8 vec
    int32_t perspective;          // 1: (me, opponent) order of the observation tensors
    const uint8_t* me;            // per-row perspective player for plain observe
    int32_t* board_of;            // KF_EXPAND
    uint8_t* keys_out;
    uint8_t* key_len_out;
    int32_t lmax_out;
    unsigned long long* stats;    // [STAT_SLICES][STAT_WORDS]
    int32_t np, nv;               // players per game and uint4 vectors per game (gamen_kernel; the two-player kernels know both at compile time)
    DevTabs tabs;
};

// Shared memory of the game kernel: one private region per warp, every array laid out [index][lane] (stride 32 words:
// bank = lane, conflict-free for any per-thread index).  The output staging buffers reuse the warp's column area, which is
// dead once the state has been written back, so only __syncwarp() is ever needed.
constexpr int WS = 32;   // stride of all per-thread arrays
template <class G> struct SmemLayout {
    typedef typename G::col_t col_t;
    static constexpr int HW = G::H * G::W;
    static constexpr int WB = 2 * HW;                         // words of packed field bits per warp (32 rows x 2HW bits)
    static constexpr int COLS = 2 * G::NCOL * WS * (int)sizeof(col_t);
    static constexpr int STREAM = 2 * G::SW * WS * 4;
    static constexpr int KEYS = G::KEYW * WS * 4;             // the acting player's packed key string
    static constexpr int GQ = 2 * GQ_WORDS * WS * 4;          // both players' garbage queues
    static constexpr int BITS = (((WB + 1) * 4 + 15) / 16) * 16;   // staging (aliases COLS): packed field bits of 32 rows
    static constexpr int VEC = 32 * 2 * VEC_LEN;                    // staging (aliases COLS): 32 rows x 38 bytes
    static_assert(BITS + VEC <= COLS, "output staging must fit in the warp's column area");
    static_assert(32 * 2 * G::SW * 4 <= BITS + VEC, "B2T_BITS staging (32 rows x 2*SW words) must fit in the staging area");
    static constexpr int O_STREAM = COLS, O_KEYS = O_STREAM + STREAM, O_GQ = O_KEYS + KEYS;
    static constexpr int WARP_BYTES = ((O_GQ + GQ + 15) / 16) * 16;
    static constexpr int TAB = 64 * 2;
    static constexpr int GAME_TOTAL = WARPS * WARP_BYTES + TAB;
};

// ---------------------------------------------------------------------------------------------- cooperative outputs
// wb: this warp's WB(+1) words.  Rows = lanes.  out = address of the warp's first row.
// B = bytes per output cell: 1 (uint8, the reference's state_dict dtype) or 2 / 4 (half / bfloat16 / float written directly for
// the network, agents/agent_utils/state_unpack.py:88-105); `one` = the bit pattern of 1.0 in that type.
template <int B> __device__ __forceinline__ uint4 expand_cells(uint32_t bits, uint32_t one) {
    uint4 q;
    if constexpr (B == 1) {          // 16 cells
        q.x = ((bits & 15u) * 0x00204081u) & 0x01010101u;
        q.y = (((bits >> 4) & 15u) * 0x00204081u) & 0x01010101u;
        q.z = (((bits >> 8) & 15u) * 0x00204081u) & 0x01010101u;
        q.w = ((bits >> 12) * 0x00204081u) & 0x01010101u;
    } else if constexpr (B == 2) {   // 8 cells, two per word
        q.x = ((bits & 1u) | ((bits & 2u) << 15)) * one;
        q.y = (((bits >> 2) & 1u) | ((bits & 8u) << 13)) * one;
        q.z = (((bits >> 4) & 1u) | ((bits & 32u) << 11)) * one;
        q.w = (((bits >> 6) & 1u) | ((bits & 128u) << 9)) * one;
    } else {                         // 4 cells
        q.x = (bits & 1u) ? one : 0u; q.y = (bits & 2u) ? one : 0u; q.z = (bits & 4u) ? one : 0u; q.w = (bits & 8u) ? one : 0u;
    }
    return q;
}
template <class G, int B = 1>
__device__ __forceinline__ void emit_fields(uint32_t* wb, const Board<G, WS>& b0, const Board<G, WS>& b1, bool has_row,
                                            uint8_t* out, int valid_rows, int lane, uint32_t one = 1u) {
    constexpr int HW = G::H * G::W;
    constexpr int REM = HW & 31;                 // valid bits in a player's last stream word (0 = all 32)
    constexpr int RW = (2 * HW + 31) / 32;       // words of one row's record (player 0 bits, then player 1 bits)
    // 1. this lane's 2*HW-bit record in registers
    uint32_t rec[RW + 1];
#pragma unroll
    for (int j = 0; j <= RW; j++) rec[j] = 0u;
    if (has_row) {
#pragma unroll
        for (int j = 0; j < G::SW; j++) rec[j] = b0.sw(j);
#pragma unroll
        for (int j = 0; j < G::SW; j++) {
            const uint32_t v = b1.sw(j);
            if (REM == 0) rec[G::SW + j] = v;
            else {
                rec[G::SW - 1 + j] |= v << REM;
                if (G::SW + j <= RW) rec[G::SW + j] |= v >> (32 - REM);
            }
        }
    }
    __syncwarp();   // the staging words alias the warp's columns: every lane must be done with its board
    // 2. place it at bit lane*2*HW of the warp's stream: shift by sh, first/last words are shared with the neighbours
    const int bit0 = lane * 2 * HW, wo = bit0 >> 5, sh = bit0 & 31;
    const int nbit0 = bit0 + 2 * HW, nsh = nbit0 & 31, nidx = (nbit0 >> 5) - wo;   // next lane's first word, relative
    uint32_t o[RW + 1];
#pragma unroll
    for (int j = 0; j <= RW; j++) o[j] = __funnelshift_l(j ? rec[j - 1] : 0u, rec[j], sh);   // (rec[j]:rec[j-1]) << sh, high word
    const uint32_t from_next = __shfl_down_sync(0xFFFFFFFFu, o[0], 1);
    const int first = sh ? 1 : 0;                       // my word 0 is written by the previous lane when shared
    const int lastk = nidx - (nsh ? 0 : 1);
#pragma unroll
    for (int j = 0; j <= RW; j++) {
        uint32_t v = o[j];
        if (j == nidx && nsh && lane < 31) v |= from_next;
        if (j >= first && j <= lastk) wb[wo + j] = v;
    }
    __syncwarp();
    // 3. every lane expands 16 / 8 / 4 bits -> 16 bytes -> one 128-bit store
    if constexpr (B != 1) {
        constexpr int CPL = 16 / B;                // cells per store
        constexpr int NCHW = 64 * HW / CPL;        // stores for 32 rows
        const int valid_ch = valid_rows * 2 * HW / CPL;   // 2*HW*B is a multiple of 16 for B >= 2 whenever HW is even ...
        static_assert((2 * G::H * G::W * B) % 16 == 0 || B == 1, "wide cells need 16-byte rows");
        const uint8_t* bb = reinterpret_cast<const uint8_t*>(wb);
        uint4* op = reinterpret_cast<uint4*>(out);
        for (int c = lane; c < NCHW && c < valid_ch; c += 32) {
            const uint32_t bits = (B == 2) ? bb[c] : ((bb[c >> 1] >> (4 * (c & 1))) & 15u);
            op[c] = expand_cells<B>(bits, one);
        }
        __syncwarp();
        return;
    }
    const uint16_t* hb = reinterpret_cast<const uint16_t*>(wb);
    constexpr int NCH = 4 * HW;  // 16-byte chunks in 32 rows
    if (valid_rows == 32) {
        const uint16_t* hp = hb + lane;
        uint4* op = reinterpret_cast<uint4*>(out) + lane;
#pragma unroll 4
        for (int it = 0; it < NCH / 32; it++) op[it * 32] = expand_cells<1>(hp[it * 32], 1u);
        if ((NCH & 31) != 0 && lane < (NCH & 31)) op[(NCH / 32) * 32] = expand_cells<1>(hp[(NCH / 32) * 32], 1u);
    } else {
        const int valid_bytes = valid_rows * 2 * HW;
        for (int c = lane; c < NCH; c += 32) {
            const int byte0 = 16 * c;
            if (byte0 >= valid_bytes) break;
            const uint4 q = expand_cells<1>(hb[c], 1u);
            if (byte0 + 16 <= valid_bytes) {
                *reinterpret_cast<uint4*>(out + byte0) = q;
            } else {
#pragma unroll
                for (int kk = 0; kk < 16; kk++) {
                    const uint32_t word = (kk < 4) ? q.x : (kk < 8) ? q.y : (kk < 12) ? q.z : q.w;
                    if (byte0 + kk < valid_bytes) out[byte0 + kk] = (uint8_t)(word >> (8 * (kk & 3)));
                }
            }
        }
    }
    __syncwarp();
}

// B2T_BITS: the fields as they are kept — one bit per cell, bit row*W+col of a player's SW words — [rows, 2, SW] uint32.  An extra
// format for consumers that unpack on the device: 8x fewer bytes than the reference's uint8 cells and no expansion work.
template <class G>
__device__ __forceinline__ void emit_field_bits(uint32_t* wb, const Board<G, WS>& b0, const Board<G, WS>& b1, bool has_row, uint32_t* out,
                                                int valid_rows, int lane) {
    constexpr int RW2 = 2 * G::SW;
    uint32_t rec[RW2];
#pragma unroll
    for (int j = 0; j < G::SW; j++) { rec[j] = has_row ? b0.sw(j) : 0u; rec[G::SW + j] = has_row ? b1.sw(j) : 0u; }
    __syncwarp();   // the staging words alias the warp's columns: every lane must be done with its board
#pragma unroll
    for (int j = 0; j < RW2; j++) wb[lane * RW2 + j] = rec[j];
    __syncwarp();
    const int total = valid_rows * RW2;                    // words
    const uint4* s4 = reinterpret_cast<const uint4*>(wb);
    uint4* o4 = reinterpret_cast<uint4*>(out);
    for (int c = lane; 4 * c < total; c += 32) {
        if (4 * c + 4 <= total) o4[c] = s4[c];
        else
            for (int k = 4 * c; k < total; k++) out[k] = wb[k];
    }
    __syncwarp();
}

// state_dict vector part (state_processors.py:33-41) of one player as 19 bytes in 5 words, built arithmetically:
// [x y inc ct] [cc n0 n1 n2] [n3..n6] [k0..k3] [k4 k5 k6]
__device__ __forceinline__ void vec_words(const PS& s, uint32_t* w) {
    const uint32_t ct = (uint32_t)(s.remaining + 50) & 0xFFFFu;
    const uint32_t ctp = (ct < 25000u ? ct : 25000u) / 100u;
    const uint64_t ohn = (s.next < 7) ? (1ull << (8 * s.next)) : 0ull;   // one-hot bytes
    const uint64_t ohk = (s.kind < 7) ? (1ull << (8 * s.kind)) : 0ull;
    w[0] = (uint32_t)(s.x & 255) | ((uint32_t)(s.y & 255) << 8) | ((uint32_t)(s.inc_count & 255) << 16) | (ctp << 24);
    w[1] = (uint32_t)(s.combo_count & 255) | ((uint32_t)ohn << 8);
    w[2] = (uint32_t)(ohn >> 24);
    w[3] = (uint32_t)ohk;
    w[4] = (uint32_t)(ohk >> 32);
}
// DT: 0 uint8, 1 half, 2 bfloat16, 3 float (b2t_obs.dtype)
template <int DT> __device__ __forceinline__ uint32_t cvt_small(uint32_t v) {   // exact: v <= 255
    if constexpr (DT == 1) return (uint32_t)__half_as_ushort(__uint2half_rn(v));
    else if constexpr (DT == 2) return __float_as_uint((float)v) >> 16;
    else return __float_as_uint((float)v);
}
template <int DT = 0>
__device__ __forceinline__ void emit_vecs(uint8_t* sv, const PS& p0, const PS& p1, bool swap, bool has_row, uint8_t* out, int valid_rows, int lane) {
    __syncwarp();   // staging aliases the warp's columns
    if (has_row) {
        uint32_t a[5], b[5], r[10];
        vec_words(p0, a);
        vec_words(p1, b);
        if (swap) {
#pragma unroll
            for (int j = 0; j < 5; j++) { const uint32_t t = a[j]; a[j] = b[j]; b[j] = t; }
        }
        // 38-byte record = a (19 bytes) ++ b (19 bytes): b starts 3 bytes into word 4
        r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = a[3];
        r[4] = (a[4] & 0xFFFFFFu) | (b[0] << 24);
        r[5] = (b[0] >> 8) | (b[1] << 24);
        r[6] = (b[1] >> 8) | (b[2] << 24);
        r[7] = (b[2] >> 8) | (b[3] << 24);
        r[8] = (b[3] >> 8) | (b[4] << 24);
        r[9] = (b[4] >> 8) & 0xFFFFu;
        uint16_t* d16 = reinterpret_cast<uint16_t*>(sv + lane * 2 * VEC_LEN);   // 2-byte aligned (38*lane)
#pragma unroll
        for (int k = 0; k < 19; k++) d16[k] = (uint16_t)(r[k >> 1] >> (16 * (k & 1)));
    } else if (lane < valid_rows) {   // an inert row inside the valid range (game id outside the handle): zeros, not stale staging bytes
        uint16_t* d16 = reinterpret_cast<uint16_t*>(sv + lane * 2 * VEC_LEN);
#pragma unroll
        for (int k = 0; k < 19; k++) d16[k] = 0;
    }
    __syncwarp();
    if constexpr (DT != 0) {     // the staged bytes widened on the way out: 8 (half / bfloat16) or 4 (float) values per 128-bit store
        constexpr int B = (DT == 3) ? 4 : 2, CPL = 16 / B;
        const int nval = valid_rows * 2 * VEC_LEN;
        for (int c = lane; c * CPL < nval; c += 32) {
            uint32_t v[CPL];
#pragma unroll
            for (int j = 0; j < CPL; j++) v[j] = cvt_small<DT>(sv[c * CPL + j]);   // (reads past the last row stay inside the staging area; the stores are guarded)
            if ((c + 1) * CPL <= nval) {
                uint4 q;
                if constexpr (B == 4) { q.x = v[0]; q.y = v[1]; q.z = v[2]; q.w = v[3]; }
                else { q.x = v[0] | (v[1] << 16); q.y = v[2] | (v[3] << 16); q.z = v[4] | (v[5] << 16); q.w = v[6] | (v[7] << 16); }
                reinterpret_cast<uint4*>(out)[c] = q;
            } else {
#pragma unroll
                for (int j = 0; j < CPL; j++)
                    if (c * CPL + j < nval) {
                        if constexpr (B == 4) reinterpret_cast<uint32_t*>(out)[c * CPL + j] = v[j];
                        else reinterpret_cast<uint16_t*>(out)[c * CPL + j] = (uint16_t)v[j];
                    }
            }
        }
        __syncwarp();
        return;
    }
    if (valid_rows == 32) {
        const uint4* s4 = reinterpret_cast<const uint4*>(sv);
        uint4* o4 = reinterpret_cast<uint4*>(out);
        for (int k = lane; k < 32 * 2 * VEC_LEN / 16; k += 32) o4[k] = s4[k];
    } else {
        const uint16_t* s16 = reinterpret_cast<const uint16_t*>(sv);
        uint16_t* o16 = reinterpret_cast<uint16_t*>(out);
        const int valid = valid_rows * VEC_LEN;  // halfwords
        for (int k = lane; k < valid; k += 32) o16[k] = s16[k];
    }
    __syncwarp();
}

// 16 u8 key codes -> 16 packed 4-bit tokens (codes > 10 become the no-op 0, PythonHandle.cpp:109-110)
__device__ __forceinline__ uint32_t squeeze_keys(uint32_t w) {
    w &= ~__vcmpgtu4(w, 0x0A0A0A0Au);
    uint32_t x = w & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}

// 4 packed tokens (16 bits) -> 4 key bytes
__device__ __forceinline__ uint32_t spread_nibbles(uint32_t x) {
    return (x & 0xFu) | ((x & 0xF0u) << 4) | ((x & 0xF00u) << 8) | ((x & 0xF000u) << 12);
}

// ---------------------------------------------------------------------------------------------- the game kernel
// MODE: 0 = every flag is read at run time (reset / observe / pokes / steps without write-back);  1 = expand and 2 = the plain
// stored step are compiled with their flag set fixed, so the two hot variants carry no code for the others' paths.
constexpr int GM_GENERIC = 0, GM_EXPAND = 1, GM_STEP = 2;
// (expand is latency-bound: a fifth resident block — 96 registers, no spills — is 3.5 % faster; the stored step is HBM-bound and
// loses 14 % to the same cap, measured)
template <class G, int MODE>
__global__ void __launch_bounds__(TPB, MODE == GM_EXPAND ? B2T_GAME_MINB + 1 : B2T_GAME_MINB) game_kernel(const KParams p) {
    typedef typename G::col_t col_t;
    typedef SmemLayout<G> L;
    constexpr int KW = G::KEYW;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    unsigned char* wbase = smem + warp * L::WARP_BYTES;
    col_t* s_cols = reinterpret_cast<col_t*>(wbase);
    uint32_t* s_stream = reinterpret_cast<uint32_t*>(wbase + L::O_STREAM);
    uint32_t* kw = reinterpret_cast<uint32_t*>(wbase + L::O_KEYS) + lane;   // kw[j * WS]
    uint32_t* s_gq = reinterpret_cast<uint32_t*>(wbase + L::O_GQ);
    uint32_t* s_bits = reinterpret_cast<uint32_t*>(wbase);                  // staging, valid once the columns are dead
    uint8_t* s_vec = wbase + L::BITS;
    uint16_t* s_tab = reinterpret_cast<uint16_t*>(smem + WARPS * L::WARP_BYTES);
    const int row = blockIdx.x * TPB + tid;
    const uint32_t flags = MODE == GM_EXPAND ? (KF_EXPAND | (p.flags & KF_FINALIZE))
                         : MODE == GM_STEP ? (KF_STEP | KF_STORE | (p.flags & KF_FINALIZE)) : p.flags;
    if (MODE == GM_EXPAND) {   // the grid covers the caller's capacity; blocks past the actual row count leave at once
        int t = p.offsets[p.n];
        if (t > p.capacity) t = p.capacity;
        if (blockIdx.x * TPB >= t) return;
    }
    if (tid < 28) { s_tab[tid] = p.tabs.pcn[tid]; s_tab[32 + tid] = p.tabs.pcc[tid]; }
    Tabs tb;
    tb.pcn = s_tab; tb.pcc = s_tab + 32;
    tb.piece = p.tabs.piece; tb.hole = p.tabs.hole; tb.powtab = p.tabs.powtab;
    tb.piece_len = p.tabs.piece_len; tb.hole_len = p.tabs.hole_len; tb.spawn_rots = p.tabs.spawn_rots;

    Game<G, WS> g;
#pragma unroll
    for (int pl = 0; pl < 2; pl++) {
        g.b[pl].cols = s_cols + pl * G::NCOL * WS + lane;
        g.b[pl].stream = s_stream + pl * G::SW * WS + lane;
        g.p[pl].gq = s_gq + pl * GQ_WORDS * WS + lane;
        board_init_walls<G, WS>(g.b[pl]);
    }
    __syncthreads();

    const bool expand = (flags & KF_EXPAND) != 0;
    const bool stepping = (flags & (KF_STEP | KF_EXPAND)) != 0;
    int total_rows = p.n, i = row, k = 0;
    if (expand) {
        total_rows = p.offsets[p.n];
        if (total_rows > p.capacity) total_rows = p.capacity;
    }
    // per-thread stat contributions
    int st_steps = 0, st_done = 0, st_w0 = 0, st_w1 = 0, st_both = 0, st_lines = 0, st_gq = 0, st_rng = 0;
    int obs_player = 0;
    int gi = 0, player = 0, sel = 0, nkeys = 0;
    bool active = row < total_rows;
    if (active) {
        if (expand) {
            const int4 m = p.row_meta[row];
            i = m.x; gi = m.y; k = m.z; player = m.w ? 1 : 0; sel = k;
        } else gi = p.idx ? p.idx[i] : i + p.game0;
        // a game id outside the handle (the caller's list is never trusted with the state array, the placement cache or kcache):
        // the row is inert — nothing is read or written for it and its observation rows come out as zeros
        active = (unsigned)gi < (unsigned)p.n_games;
    }

    if (active) {
        // ---- every global load that does not depend on the state is issued before the state itself
        if (!expand) {
            if (stepping) {
                player = p.player[i] ? 1 : 0;
                if (p.placement) sel = p.placement[i]; else nkeys = (MODE != GM_GENERIC || p.key_len) ? (int)p.key_len[i] : -1;
            }
        }
        obs_player = player;
        const bool by_cache = stepping && (expand || p.placement);
        // fast path: key rows of 16 or 32 bytes, 16-byte aligned, fetched as one or two 128-bit loads
        constexpr int KQ = 2;
        const bool key_vec = stepping && !by_cache && (MODE != GM_GENERIC || p.key_fmt == 0) && (p.lmax == 16 || p.lmax == 32) && ((reinterpret_cast<uintptr_t>(p.keys) & 15) == 0);
        // (issuing the key loads here, ahead of the state, or prefetching the piece-stream entries was measured slower: the extra
        // live registers cost more than the exposed latency)
        int kc = 0;
        if (by_cache) kc = p.kcache[gi];
        {
            const uint4* src = p.state + state_index<G>(gi, 0);
#pragma unroll
            for (int v = 0; v < G::NV; v++) {
                const uint4 q = src[v * 32];
                game_put_word(g, 4 * v + 0, q.x);
                game_put_word(g, 4 * v + 1, q.y);
                game_put_word(g, 4 * v + 2, q.z);
                game_put_word(g, 4 * v + 3, q.w);
            }
        }
        game_unpack(g);
        if (MODE == GM_GENERIC && p.tiles && !expand) {   // the planes are updated in place in HBM (same tile-of-32 layout, stride 32)
            uint32_t* tp = p.tiles + (size_t)(gi >> 5) * (2 * TILE_PLANES * G::SW * 32) + (gi & 31);
            g.b[0].tiles = tp;
            g.b[1].tiles = tp + TILE_PLANES * G::SW * 32;
        }
        const int flags_before = g.p[0].flags | g.p[1].flags;
        if (flags & KF_RESET) game_reset(g, tb, p.seeds[i]);
        if (flags & KF_POKE_DEAD) {
            if (p.poke_player == 0) g.p[0].dead = p.poke_value; else g.p[1].dead = p.poke_value;
        }
        if (flags & KF_CLEAR_FIELD) {
            if (p.poke_player == 0) board_clear<G, WS>(g.b[0]); else board_clear<G, WS>(g.b[1]);
        }
        if (stepping) {
            if (by_cache) {
                const bool valid = sel >= 0 && sel < kc && sel < G::KMAX;   // out of range -> null action (action_list.py:27-33)
                const uint4* src = reinterpret_cast<const uint4*>(p.plc + ((size_t)gi * G::KMAX + (valid ? sel : 0)) * KW);
                nkeys = 1;
#pragma unroll
                for (int j = 0; j < KW / 4; j++) {
                    uint4 q = src[j];
                    if (!valid) q = make_uint4(0u, 0u, 0u, 0u);
                    kw[(4 * j) * WS] = q.x; kw[(4 * j + 1) * WS] = q.y; kw[(4 * j + 2) * WS] = q.z; kw[(4 * j + 3) * WS] = q.w;
                    // sequences hold no zero token (except the null action): length = position of the last non-zero nibble
                    if (q.x) nkeys = (4 * j) * 8 + (msb32(q.x) >> 2) + 1;
                    if (q.y) nkeys = (4 * j + 1) * 8 + (msb32(q.y) >> 2) + 1;
                    if (q.z) nkeys = (4 * j + 2) * 8 + (msb32(q.z) >> 2) + 1;
                    if (q.w) nkeys = (4 * j + 3) * 8 + (msb32(q.w) >> 2) + 1;
                }
            } else {
                const bool zterm = MODE == GM_GENERIC && nkeys < 0;            // zero-terminated row: every token counts, the length follows from the words
                if (zterm || nkeys > KW * 8) nkeys = KW * 8;
                if (nkeys > p.lmax) nkeys = p.lmax;
                if (MODE == GM_GENERIC && p.key_fmt == 1) {   // packed tokens: the row already is the working form (generic variant only)
                    const uint8_t* kp = p.keys + (size_t)i * (size_t)(p.lmax >> 1);
                    uint32_t w[KW];
                    const bool al = (reinterpret_cast<uintptr_t>(p.keys) & 15) == 0;
                    if (p.lmax == 16 && al) {
                        const uint2 q = *reinterpret_cast<const uint2*>(kp);
#pragma unroll
                        for (int j = 0; j < KW; j++) w[j] = j == 0 ? q.x : j == 1 ? q.y : 0u;
                    } else if (p.lmax == 32 && al && KW >= 4) {
                        const uint4 q = *reinterpret_cast<const uint4*>(kp);
#pragma unroll
                        for (int j = 0; j < KW; j++) w[j] = j == 0 ? q.x : j == 1 ? q.y : j == 2 ? q.z : j == 3 ? q.w : 0u;
                    } else {
#pragma unroll
                        for (int j = 0; j < KW; j++) {
                            uint32_t v = 0;
#pragma unroll
                            for (int b = 0; b < 4; b++)
                                if (8 * j + 2 * b < p.lmax) v |= (uint32_t)kp[4 * j + b] << (8 * b);
                            w[j] = v;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < KW; j++) kw[j * WS] = w[j];
                } else if (key_vec) {
                    uint4 kq[KQ];
                    {
                        const uint8_t* kp = p.keys + (size_t)i * (size_t)p.lmax;
#pragma unroll
                        for (int j = 0; j < KQ; j++)
                            kq[j] = (16 * j < p.lmax) ? *reinterpret_cast<const uint4*>(kp + 16 * j) : make_uint4(0u, 0u, 0u, 0u);
                    }
#pragma unroll
                    for (int j = 0; j < KW / 2; j++) {       // 16 keys per uint4 -> two packed words
                        const bool in = j < KQ && 16 * j < nkeys;
                        const uint4 q = kq[j < KQ ? j : 0];
                        kw[(2 * j) * WS] = in ? (squeeze_keys(q.x) | (squeeze_keys(q.y) << 16)) : 0u;
                        kw[(2 * j + 1) * WS] = in ? (squeeze_keys(q.z) | (squeeze_keys(q.w) << 16)) : 0u;
                    }
                } else {
                    const uint8_t* kp = p.keys + (size_t)i * (size_t)p.lmax;
                    uint32_t tmp[KW];
                    pack_keys<KW>(kp, nkeys, tmp);
#pragma unroll
                    for (int j = 0; j < KW; j++) kw[j * WS] = tmp[j];
                }
                if (zterm) {
                    nkeys = 1;
#pragma unroll
                    for (int j = 0; j < KW; j++) {
                        const uint32_t w = kw[j * WS];
                        if (w) nkeys = j * 8 + (msb32(w) >> 2) + 1;
                    }
                    if (nkeys > p.lmax) nkeys = p.lmax;
                }
            }
            const bool was_over = (g.gf() & GF_ROUND_OVER) != 0;
            const bool finalize = (flags & KF_FINALIZE) != 0;
            int rew;
            const bool done = game_perform<G, WS, WS>(g, tb, player, kw, nkeys, finalize, p.elapsed, &rew);
            if (p.reward) p.reward[row] = (int8_t)rew;
            if (p.done) p.done[row] = done ? 1 : 0;
            if (p.lines) *reinterpret_cast<uint16_t*>(p.lines + 2 * (size_t)row) = (uint16_t)((g.p[0].reward & 255) | ((g.p[1].reward & 255) << 8));
            if (expand) {
                if (p.key_len_out) p.key_len_out[row] = (uint8_t)nkeys;
                if (p.keys_out) {
                    uint8_t* ko = p.keys_out + (size_t)row * (size_t)p.lmax_out;
                    const bool vec_ok = (p.lmax_out & 15) == 0 && (reinterpret_cast<uintptr_t>(p.keys_out) & 15) == 0;
#pragma unroll
                    for (int c = 0; c < KW / 2; c++) {          // 16 keys (two packed words) per 128-bit store
                        if (16 * c < p.lmax_out) {
                            uint4 q;
                            const uint32_t k0 = kw[(2 * c) * WS], k1 = kw[(2 * c + 1) * WS];
                            q.x = spread_nibbles(k0 & 0xFFFFu); q.y = spread_nibbles(k0 >> 16);
                            q.z = spread_nibbles(k1 & 0xFFFFu); q.w = spread_nibbles(k1 >> 16);
                            if (vec_ok) *reinterpret_cast<uint4*>(ko + 16 * c) = q;
                            else {
#pragma unroll
                                for (int b = 0; b < 16; b++) {
                                    const uint32_t word = (b < 4) ? q.x : (b < 8) ? q.y : (b < 12) ? q.z : q.w;
                                    if (16 * c + b < p.lmax_out) ko[16 * c + b] = (uint8_t)(word >> (8 * (b & 3)));
                                }
                            }
                        }
                    }
                    for (int j = KW * 8; j < p.lmax_out; j++) ko[j] = 0;
                }
            } else if (finalize && !was_over) {
                st_steps = 1;
                st_lines = g.p[0].reward + g.p[1].reward;
                if (done) {
                    st_done = 1;
                    const int d0 = g.p[0].dead, d1 = g.p[1].dead;
                    st_both = d0 && d1; st_w0 = !d0 && d1; st_w1 = d0 && !d1;
                }
            }
            if (!expand && done && (p.reset_seeds || (MODE == GM_GENERIC && p.auto_reset)))
                game_reset(g, tb, p.reset_seeds ? p.reset_seeds[i] : auto_seed(p.seed_base, (uint32_t)gi, g.p[0].episode));
        }
        const int flags_after = g.p[0].flags | g.p[1].flags;
        st_gq = ((flags_after & ~flags_before) & 1) != 0;
        st_rng = ((flags_after & ~flags_before) & 2) != 0;
        if ((flags & KF_STORE) && !expand) {
            game_pack(g);
            uint4* dst = p.state + state_index<G>(gi, 0);
#pragma unroll
            for (int v = 0; v < G::NV; v++) {
                uint4 q;
                q.x = game_get_word(g, 4 * v + 0);
                q.y = game_get_word(g, 4 * v + 1);
                q.z = game_get_word(g, 4 * v + 2);
                q.w = game_get_word(g, 4 * v + 3);
                dst[v * 32] = q;
            }
        }
        if (p.piece_idx) {
            const bool sw = p.perspective && (stepping ? (player != 0) : (p.me[row] != 0));
            const int k0 = sw ? g.p[1].kind : g.p[0].kind, k1 = sw ? g.p[0].kind : g.p[1].kind;
            *reinterpret_cast<uint16_t*>(p.piece_idx + 2 * (size_t)row) = (uint16_t)((k0 & 255) | ((k1 & 255) << 8));
        }
        if (p.scalars) {
            int32_t* s = p.scalars + (size_t)row * 16;
#pragma unroll
            for (int pl = 0; pl < 2; pl++) {
                const PS& q = g.p[pl];
                int4 a = make_int4(q.x, q.y, q.inc_count, q.remaining), b = make_int4(q.combo_count, q.next, q.reward, q.dead);
                reinterpret_cast<int4*>(s)[2 * pl] = a;
                reinterpret_cast<int4*>(s)[2 * pl + 1] = b;
            }
        }
    }
    // ---- warp-cooperative observation tensors
    bool swap = false;     // perspective order: the acting player first
    if (p.perspective && active) swap = (stepping ? (obs_player != 0) : (p.me[row] != 0));
    const int warp_row0 = blockIdx.x * TPB + warp * 32;
    int valid_rows = total_rows - warp_row0;
    valid_rows = valid_rows < 0 ? 0 : (valid_rows > 32 ? 32 : valid_rows);
    const Board<G, WS>& fb0 = swap ? g.b[1] : g.b[0];
    const Board<G, WS>& fb1 = swap ? g.b[0] : g.b[1];
    if (p.obs_dtype == 4) {                            // fields as bits (B2T_BITS), vector part as uint8
        if (p.field && valid_rows > 0) emit_field_bits<G>(s_bits, fb0, fb1, active, reinterpret_cast<uint32_t*>(p.field) + (size_t)warp_row0 * (size_t)(2 * G::SW), valid_rows, lane);
        if (p.vec && valid_rows > 0) emit_vecs(s_vec, g.p[0], g.p[1], swap, active, p.vec + (size_t)warp_row0 * (size_t)(2 * VEC_LEN), valid_rows, lane);
    } else if (MODE != GM_GENERIC || p.obs_dtype == 0) {      // uint8: the reference's dtype; with B2T_BITS the only ones the two hot variants carry
        if (p.field && valid_rows > 0) emit_fields<G>(s_bits, fb0, fb1, active, p.field + (size_t)warp_row0 * (size_t)(2 * L::HW), valid_rows, lane);
        if (p.vec && valid_rows > 0) emit_vecs(s_vec, g.p[0], g.p[1], swap, active, p.vec + (size_t)warp_row0 * (size_t)(2 * VEC_LEN), valid_rows, lane);
    } else if (p.obs_dtype == 3) {                     // float
        if (p.field && valid_rows > 0) emit_fields<G, 4>(s_bits, fb0, fb1, active, p.field + (size_t)warp_row0 * (size_t)(8 * L::HW), valid_rows, lane, 0x3F800000u);
        if (p.vec && valid_rows > 0) emit_vecs<3>(s_vec, g.p[0], g.p[1], swap, active, p.vec + (size_t)warp_row0 * (size_t)(8 * VEC_LEN), valid_rows, lane);
    } else {                                           // half / bfloat16
        const uint32_t one = p.obs_dtype == 1 ? 0x3C00u : 0x3F80u;
        if (p.field && valid_rows > 0) emit_fields<G, 2>(s_bits, fb0, fb1, active, p.field + (size_t)warp_row0 * (size_t)(4 * L::HW), valid_rows, lane, one);
        if (p.vec && valid_rows > 0) {
            if (p.obs_dtype == 1) emit_vecs<1>(s_vec, g.p[0], g.p[1], swap, active, p.vec + (size_t)warp_row0 * (size_t)(4 * VEC_LEN), valid_rows, lane);
            else emit_vecs<2>(s_vec, g.p[0], g.p[1], swap, active, p.vec + (size_t)warp_row0 * (size_t)(4 * VEC_LEN), valid_rows, lane);
        }
    }
    // ---- episode statistics: one atomic per non-zero counter per warp, spread over slices
    if (p.stats && (flags & KF_STEP)) {
        // the eight per-lane contributions are tiny (flags, or <= 8 lines): two packed warp reductions instead of eight
        const uint32_t pa = (uint32_t)st_steps | ((uint32_t)st_done << 6) | ((uint32_t)st_w0 << 12) | ((uint32_t)st_w1 << 18) | ((uint32_t)st_both << 24);
        const uint32_t pb = (uint32_t)st_lines | ((uint32_t)st_gq << 10) | ((uint32_t)st_rng << 16);
        const uint32_t ra = __reduce_add_sync(0xFFFFFFFFu, pa), rb = __reduce_add_sync(0xFFFFFFFFu, pb);
        if (lane == 0) {
            const uint32_t vals[STAT_WORDS] = {ra & 63u, (ra >> 6) & 63u, (ra >> 12) & 63u, (ra >> 18) & 63u, (ra >> 24) & 63u,
                                               rb & 1023u, (rb >> 10) & 63u, (rb >> 16) & 63u};
#pragma unroll
            for (int j = 0; j < STAT_WORDS; j++)
                if (vals[j]) atomicAdd(&p.stats[(blockIdx.x % STAT_SLICES) * STAT_WORDS + j], (unsigned long long)vals[j]);
        }
    }
}

// ---------------------------------------------------------------------------------------------- enumerate
struct EParams {
    const uint4* state;
    int32_t n_games;
    const int32_t* idx;
    int32_t n;
    const uint8_t* player;
    int32_t bar_null;
    uint32_t* plc;      // [n_games][KMAX][KEYW]
    int32_t* kcache;    // [n_games]
    int32_t* K;         // [n] (may be NULL)
    unsigned long long* stats;
    int32_t nv, np;     // uint4 vectors per game (G::NV with two players), players per game
    DevTabs tabs;
};

// Placement enumeration, one board per thread, in three phases so that the rare, expensive under-overhang search never
// runs with one lane active while 31 wait:
//   1. (converged) every lane scans (column, rotation) combos of ITS board: one collision mask per combo gives the top
//      drop (a bit in `topmask`) and the resting heights of the cavities below it (candidates, in shared memory); the scan
//      pauses (warp-uniformly) only if some lane's candidate slots could overflow;
//   2. (pooled) the warp's candidates form a work queue; a lane takes the next candidate as soon as its search ends and all
//      busy lanes execute Enumerator::search_step() together on the OWNERS' boards (shared memory); results return through
//      a per-warp buffer and each owner accepts its results in candidate order (first-occurrence de-duplication,
//      action_list.py:14) straight into their final list slots:
//      slot = #top drops with combo <= c  +  #finesse entries accepted before;
//   3. every lane writes its top drops into the remaining slots.
// Emission order is exactly TestField::getMask's: combos ascending (x-major, rotation-minor), top drop first, then the
// combo's cavities by increasing height.
constexpr int EWARPS = 2;             // warps per block that OWN boards (small: 65 536 boards must still fill 148 SMs)
#ifndef B2T_ENUM_HELPERS
#define B2T_ENUM_HELPERS 1
#endif
// Helper warps.  65 536 boards are only 2 048 owner warps — 3.5 per scheduler — and the pooled search is one long chain of dependent
// integer instructions, so the kernel ran at the speed of its instruction latencies (issue slots 36 % used).  Every owner warp
// therefore gets EHELP helper warps that carry no boards and only pull candidates from the owner's pool during phase 2 (pairwise
// named barriers around each pooled search); the registers are capped so that owners and helpers of all boards are resident at once.
constexpr int EHELP = B2T_ENUM_HELPERS;
constexpr int ETPB = 32 * EWARPS * (1 + EHELP);   // threads per block of the enumerate kernel
constexpr int EMINB = 28 / (ETPB / 32) < 1 ? 1 : 28 / (ETPB / 32);   // resident blocks the register cap is chosen for (28 warps per SM)
#ifndef B2T_ENUM_BPW
#define B2T_ENUM_BPW 32
#endif
// Boards per warp.  Phases 1 and 3 run one board per lane (lanes >= BPW idle); the pooled search of phase 2 always uses all 32
// lanes on the warp's candidates.  Measured: 16 and 8 boards per warp (more, shorter warps) execute 39 % / 89 % more
// instructions for the same boards and are slower than 32 (292 / 329 us vs 253 us for 65 536 boards).
constexpr int BPW = B2T_ENUM_BPW;
constexpr int ENB = EWARPS * BPW;     // boards per block = stride of every per-board array in shared memory

template <class G> struct EnumCfg {
    static constexpr int CCAP = 32;                                  // candidate slots per thread between two pooled searches
    static constexpr int PER_COMBO = (G::H - 1) / 2;                 // resting heights differ by >= 2 and start at 2
    static constexpr int NCOMBO = 4 * G::W;
    static constexpr int ACAP = G::KMAX;                             // accepted finesse entries per board (cannot overflow before the list does)
    static constexpr int RES = 192 * BPW / 32 < 64 ? 64 : 192 * BPW / 32;   // pooled results buffered per warp between acceptances
    static constexpr int PCAP = EnumGeo<G>::PCAP;
    static constexpr int RSW = 1 + (PCAP + 3) / 4;                   // result slot: header word + the raw search path
    static_assert(NCOMBO <= 128, "topmask is 128 bits");
    static_assert(PER_COMBO <= CCAP, "a single combo must fit the candidate slots");
    static constexpr int O_CAND = G::NCOL * ENB * (int)sizeof(typename G::col_t);
    static constexpr int O_PC4 = O_CAND + CCAP * ENB * 2;
    static constexpr int O_RES = O_PC4 + ENB * 16;
    static constexpr int O_INCL = O_RES + EWARPS * RES * RSW * 4;
    static constexpr int O_Q = O_INCL + EWARPS * 32 * 4;
    static constexpr int O_KIND = O_Q + EWARPS * 4 + ((EWARPS & 1) ? 4 : 0);
    static constexpr int O_CTL = O_KIND + ENB * 4;                   // [EWARPS][2] {entries of the group (-1: no more work), first entry} for the helpers
    static constexpr int O_TAB = O_CTL + EWARPS * 8;
    static constexpr int O_ACC = O_TAB + 128;
    static constexpr int O_PBUF = ((O_ACC + ACAP * ENB + 15) / 16) * 16;
    static constexpr int SMEM = O_PBUF + G::KEYW * ENB * 4 + 16;   // O_PBUF: per-thread scratch for the sequence being accepted
};

struct NoSink {
    template <class S> B2T_HD void top(const S&, bool) {}
    template <class S> B2T_HD void finesse(const S&, bool) {}
};

__device__ __forceinline__ int rank128(uint64_t lo, uint64_t hi, int nbits) {  // set bits among the lowest nbits (0..128)
    if (nbits <= 0) return 0;
    if (nbits >= 128) return __popcll(lo) + __popcll(hi);
    if (nbits <= 64) return __popcll(nbits == 64 ? lo : (lo & ((1ull << nbits) - 1ull)));
    return __popcll(lo) + __popcll(hi & ((1ull << (nbits - 64)) - 1ull));
}

template <class G>
__global__ void __launch_bounds__(ETPB, EMINB) enumerate_kernel(const EParams p) {
    typedef typename G::col_t col_t;
    typedef EnumCfg<G> C;
    constexpr int KW = G::KEYW;
    extern __shared__ __align__(16) unsigned char smem[];
    col_t* s_cols = reinterpret_cast<col_t*>(smem);
    uint16_t* s_cand = reinterpret_cast<uint16_t*>(smem + C::O_CAND);   // [CCAP][ETPB]
    uint64_t* s_pc4 = reinterpret_cast<uint64_t*>(smem + C::O_PC4);     // [ETPB][2] pcn4, pcc4 of the owner's piece
    uint32_t* s_res = reinterpret_cast<uint32_t*>(smem + C::O_RES);     // [EWARPS][RES][RSW] header + raw path of every pooled search
    int* s_incl = reinterpret_cast<int*>(smem + C::O_INCL);             // [EWARPS][32] inclusive candidate counts
    int* s_q = reinterpret_cast<int*>(smem + C::O_Q);                   // [EWARPS] work-queue heads
    int* s_kind = reinterpret_cast<int*>(smem + C::O_KIND);             // [ENB]
    volatile int* s_ctl = reinterpret_cast<volatile int*>(smem + C::O_CTL);
    uint16_t* s_tab = reinterpret_cast<uint16_t*>(smem + C::O_TAB);
    uint8_t* s_acc = smem + C::O_ACC;                                   // [ACAP][ETPB] combo of every accepted finesse entry
    uint32_t* s_seq = reinterpret_cast<uint32_t*>(smem + C::O_PBUF);    // [KW][ETPB] key string being accepted by this owner
    const int tid0 = threadIdx.x, lane = tid0 & 31, wid = tid0 >> 5;
    const bool helper = wid >= EWARPS;                  // helper warps own no boards: they only search
    const int warp = helper ? (wid - EWARPS) % EWARPS : wid;   // the pool this warp works on (= index of the owner warp)
    const bool has = !helper && lane < BPW;             // this lane carries a board
    const int tid = warp * BPW + (has ? lane : 0);      // board slot in the block's shared arrays (idle lanes alias slot 0, read-only)
    const int i = blockIdx.x * ENB + tid;
    if (tid0 < 28) { s_tab[tid0] = p.tabs.pcn[tid0]; s_tab[32 + tid0] = p.tabs.pcc[tid0]; }
    Tabs tb;
    tb.pcn = s_tab; tb.pcc = s_tab + 32;
    tb.piece = p.tabs.piece; tb.hole = p.tabs.hole; tb.powtab = p.tabs.powtab;
    tb.piece_len = p.tabs.piece_len; tb.hole_len = p.tabs.hole_len; tb.spawn_rots = p.tabs.spawn_rots;
    Board<G, ENB> b;
    b.cols = s_cols + tid;
    b.stream = nullptr;
    if (has) board_init_walls<G, ENB>(b);
    __syncthreads();
    bool active = has && i < p.n;
    int gi = 0, kind = 7, cur_rot = 0;
    if (active) {
        gi = p.idx ? p.idx[i] : i;
        if ((unsigned)gi >= (unsigned)p.n_games) {   // id outside the handle: an empty list, nothing touched
            active = false; gi = 0;
            if (p.K) p.K[i] = 0;
        }
    }
    if (active) {
        const int pl = p.player[i] < p.np ? p.player[i] : p.np - 1;
        const uint32_t* words = reinterpret_cast<const uint32_t*>(p.state);
        const size_t tile = (size_t)(gi >> 5) * (size_t)p.nv;      // (nv is G::NV unless the handle has another player count)
        auto word = [&](int wi) { return words[((tile + (size_t)(wi >> 2)) * 32 + (size_t)(gi & 31)) * 4 + (wi & 3)]; };
        const int base = pl * G::PW;
        for (int c = 0; c < G::W; c++) {
            if (G::CW == 1) b.col(c) = (col_t)word(base + c);
            else b.col(c) = (col_t)((uint64_t)word(base + 2 * c) | ((uint64_t)word(base + 2 * c + 1) << 32));
        }
        const uint32_t w0 = word(base + G::W * G::CW + G::SW);
        kind = w0 & 7; cur_rot = (w0 >> 3) & 3;
    }
    const int kk = kind <= 6 ? kind : 0;
    uint64_t pcn4 = 0, pcc4 = 0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
        pcn4 |= (uint64_t)tb.pcn[kk * 4 + r] << (16 * r);
        pcc4 |= (uint64_t)tb.pcc[kk * 4 + r] << (16 * r);
    }
    if (has) { s_pc4[2 * tid] = pcn4; s_pc4[2 * tid + 1] = pcc4; s_kind[tid] = kk; }
    const int nrot = (kind > 6) ? 0 : (kind == 6) ? 1 : (kind == 4 || kind == 2 || kind == 3) ? 2 : 4;
    const int spawn = spawn_rot_of(tb, kk);
    const int off = p.bar_null ? 0 : 1;
    uint32_t* lst = p.plc + (size_t)gi * G::KMAX * KW;      // this board's list (only dereferenced by active lanes)
    uint64_t top_lo = 0, top_hi = 0;
    int nacc = 0, bad = 0;
    NoSink ns;
    Enumerator<G, ENB, KW, NoSink, 1> e(b, tb, kk, ns, nullptr);
    __syncwarp();

    // the owner warp of a pool and its helpers meet at a named barrier before and after every pooled search
    auto pool_barrier = [&]() {
        static_assert(EWARPS <= 2, "one named barrier per pool: add cases");
        if (EHELP == 0) __syncwarp();
        else if (warp == 0) asm volatile("bar.sync 1, %0;" ::"n"(32 * (1 + EHELP)) : "memory");      // immediate ids: the kernel reserves 3 barriers, not 16
        else asm volatile("bar.sync 2, %0;" ::"n"(32 * (1 + EHELP)) : "memory");
    };
    // phase 2 proper: every lane pulls the pool's next candidate as soon as its search ends; all busy lanes step together
    auto pool_search = [&](int g0, int n_g) {
        bool busy = false;
        int myj = 0;
        for (;;) {
            if (!busy) {
                const int j = atomicAdd(&s_q[warp], 1);
                if (j < n_g) {
                    const int g = g0 + j;
                    int o = 0;   // owner = first lane whose inclusive count exceeds g
#pragma unroll
                    for (int step = 16; step >= 1; step >>= 1) {
                        const int t = o + step;
                        if (t <= 31 && s_incl[warp * 32 + t - 1] <= g) o = t;
                    }
                    const int o_excl = o ? s_incl[warp * 32 + o - 1] : 0;
                    const int otid = warp * BPW + o;
                    const uint32_t cd = s_cand[(g - o_excl) * ENB + otid];
                    Board<G, ENB> ob;
                    ob.cols = s_cols + otid;
                    ob.stream = nullptr;
                    const int okind = s_kind[otid];
                    e.rebind(ob, okind, spawn_rot_of(tb, okind), s_pc4[2 * otid], s_pc4[2 * otid + 1]);
                    e.pbuf = reinterpret_cast<uint8_t*>(s_res + (warp * C::RES + j) * C::RSW + 1);   // the path is built in the result slot
                    e.path_clear();
                    e.search_begin((int)(cd >> 10) - 1, (int)(cd & 63u), (int)((cd >> 6) & 3u));
                    busy = true;
                    myj = j;
                }
            }
            if (__all_sync(0xFFFFFFFFu, !busy)) break;
            if (busy && e.search_step()) {
                s_res[(warp * C::RES + myj) * C::RSW] = e.found
                    ? (0x80000000u | ((uint32_t)e.path_n << 16) | ((uint32_t)((e.best_x + 8) & 255) << 8) | ((uint32_t)e.best_rot << 2) | (e.path_bad ? 1u : 0u))
                    : 0u;
                busy = false;
            }
        }
    };
    if (helper) {            // helpers: wait for a group of the owner's candidates, search it with the owner, repeat until told to leave
        for (;;) {
            pool_barrier();
            const int n_g = s_ctl[2 * warp], g0 = s_ctl[2 * warp + 1];
            if (n_g < 0) return;
            pool_search(g0, n_g);
            pool_barrier();
        }
    }

    int c_next = 0;
    while (c_next < C::NCOMBO) {
        // ---- phase 1: my combos, until some lane could run out of candidate slots
        int ncand = 0, c = c_next;
        for (; c < C::NCOMBO; c++) {
            if (__any_sync(0xFFFFFFFFu, ncand > C::CCAP - C::PER_COMBO)) break;
            const int r0 = c & 3, xx = (c >> 2) - 1;
            if (r0 < nrot) {
                const int r = (kind == 6) ? cur_rot : r0;
                const col_t cm = collide_mask<G, ENB>(b, xx, (uint32_t)(pcc4 >> (16 * r)) & 0xFFFFu);
                if (!((cm >> G::YB) & 1)) {
                    if (c < 64) top_lo |= 1ull << c; else top_hi |= 1ull << (c - 64);
                    int yy = rest_from<G>(cm, 0) + 2;
                    while (yy < G::H - 1) {
                        col_t free_ = (col_t)(~cm >> (yy + G::YB));
                        free_ &= (col_t)((G::one() << (G::H - 1 - yy)) - 1);
                        if (!free_) break;
                        const int rest = rest_from<G>(cm, yy + ctz(free_));
                        if (ncand < C::CCAP) s_cand[ncand * ENB + tid] = (uint16_t)((c << 8) | (r << 6) | rest);
                        else bad = 1;
                        ncand++;
                        yy = rest + 1;
                    }
                }
            }
        }
        c_next = c;
        if (ncand > C::CCAP) ncand = C::CCAP;
        // ---- phase 2: pool the warp's candidates
        int incl = ncand;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        const int total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const int excl = incl - ncand;
        s_incl[warp * 32 + lane] = incl;
        __syncwarp();
        for (int g0 = 0; g0 < total; g0 += C::RES) {
            const int n_g = (total - g0) < C::RES ? (total - g0) : C::RES;
            if (lane == 0) { s_q[warp] = 0; s_ctl[2 * warp] = n_g; s_ctl[2 * warp + 1] = g0; }
            pool_barrier();          // the helpers see this group (and the candidates, counts and boards written above)
            pool_search(g0, n_g);
            pool_barrier();          // every search of the group has left its result
            // ---- collect: I am the owner of group entries [a, bnd)
            int a = excl - g0, bnd = incl - g0;
            a = a < 0 ? 0 : a; bnd = bnd > n_g ? n_g : bnd;
            // lanes first skip over their failed searches (cheap), so that the k-th SUCCESS of every owner is expanded,
            // de-duplicated and stored in the same warp iteration
            for (int t = a;; t++) {
                while (t < bnd && s_res[(warp * C::RES + t) * C::RSW] == 0u) t++;
                if (!__any_sync(0xFFFFFFFFu, t < bnd)) break;
                if (t >= bnd) continue;
                const uint32_t* slot_in = s_res + (warp * C::RES + t) * C::RSW;
                const uint32_t hd = slot_in[0];
                // expand the recorded move into its forward key string (useFinesseMove, TestField.cpp:127-164)
                uint32_t* rr = s_seq + tid;                     // word j at rr[j * ENB]
                WordWriter<KW, ENB> w(rr);
                Enumerator<G, ENB, KW, NoSink, 1>::template write_finesse_raw<WordWriter<KW, ENB>, 1>(
                    w, reinterpret_cast<const uint8_t*>(slot_in + 1), (int)((hd >> 16) & 255u), (int)((hd >> 8) & 255u) - 8, (int)((hd >> 2) & 3u), spawn);
                w.finish();
                if (w.bad || (hd & 1u)) bad = 1;
                const int cc = (int)(s_cand[(g0 + t - excl) * ENB + tid] >> 8);
                bool dup = false;
                for (int f = 0; f < nacc && !dup; f++) {
                    const int fc = s_acc[f * ENB + tid];
                    const int slot = off + rank128(top_lo, top_hi, fc + 1) + f;
                    if (slot < G::KMAX) {
                        const uint32_t* q = lst + (size_t)slot * KW;
                        bool eq = true;
#pragma unroll
                        for (int j = 0; j < KW; j++) eq = eq && (q[j] == rr[j * ENB]);
                        dup = eq;
                    }
                }
                if (dup) continue;
                const int slot = off + rank128(top_lo, top_hi, cc + 1) + nacc;
                if (slot < G::KMAX && nacc < C::ACAP) {
                    uint4* dst = reinterpret_cast<uint4*>(lst + (size_t)slot * KW);
#pragma unroll
                    for (int j = 0; j < KW / 4; j++)
                        dst[j] = make_uint4(rr[(4 * j) * ENB], rr[(4 * j + 1) * ENB], rr[(4 * j + 2) * ENB], rr[(4 * j + 3) * ENB]);
                    s_acc[nacc * ENB + tid] = (uint8_t)cc;
                } else bad = 1;
                nacc++;
            }
            __syncwarp();
        }
    }
    if (EHELP > 0) {         // no more groups: the helpers leave
        if (lane == 0) s_ctl[2 * warp] = -1;
        pool_barrier();
    }
    // ---- phase 3: top drops
    if (active) {
        if (nacc > C::ACAP) nacc = C::ACAP;
        int f = 0, ntop = 0;
        for (int c = 0; c < C::NCOMBO; c++) {
            const bool has = (c < 64) ? ((top_lo >> c) & 1ull) : ((top_hi >> (c - 64)) & 1ull);
            if (!has) continue;
            while (f < nacc && s_acc[f * ENB + tid] < c) f++;
            const int slot = off + ntop + f;
            const int r = (kind == 6) ? cur_rot : (c & 3);
            const uint64_t t = top_drop_tokens((r - spawn) & 3, ((c >> 2) - 1) - G::SPAWN_X);
            if (slot < G::KMAX) {
                uint4* dst = reinterpret_cast<uint4*>(lst + (size_t)slot * KW);
                dst[0] = make_uint4((uint32_t)t, (uint32_t)(t >> 32), 0u, 0u);
#pragma unroll
                for (int j = 1; j < KW / 4; j++) dst[j] = make_uint4(0u, 0u, 0u, 0u);
            } else bad = 1;
            ntop++;
        }
        int K = ntop + nacc + off;
        if (ntop + nacc == 0 || off) {   // null action [0] in slot 0 (action_list.py:6-12,18-20)
            uint4* dst = reinterpret_cast<uint4*>(lst);
#pragma unroll
            for (int j = 0; j < KW / 4; j++) dst[j] = make_uint4(0u, 0u, 0u, 0u);
            if (ntop + nacc == 0) K = 1;
        }
        p.kcache[gi] = K;
        if (p.K) p.K[i] = K;
        if (bad && p.stats) atomicAdd(&p.stats[(blockIdx.x % STAT_SLICES) * STAT_WORDS + 7], 1ull);  // list capacity exceeded
    }
}

// packed cache -> the u8 [n,kmax,lmax] / [n,kmax] arrays of the reference-shaped API
template <class G>
__global__ void unpack_keys_kernel(const uint32_t* plc, const int32_t* kcache, const int32_t* idx, int n_games, int n, int kmax, int lmax,
                                   uint8_t* keys, uint8_t* key_len) {
    constexpr int KW = G::KEYW;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * kmax) return;
    const int i = (int)(t / kmax), k = (int)(t % kmax);
    const int gi = idx ? idx[i] : i;
    const bool in_range = (unsigned)gi < (unsigned)n_games;
    const int K = in_range ? kcache[gi] : 0;
    uint32_t w[KW];
    const bool valid = k < K && k < G::KMAX;
    for (int j = 0; j < KW; j++) w[j] = valid ? plc[((size_t)gi * G::KMAX + k) * KW + j] : 0u;
    const int len = valid ? packed_len<KW>(w) : 0;
    if (key_len) key_len[t] = (uint8_t)(len < lmax ? len : lmax);
    if (keys)
        for (int j = 0; j < lmax; j++) keys[(size_t)t * lmax + j] = (j < KW * 8) ? (uint8_t)((w[j >> 3] >> (4 * (j & 7))) & 15u) : 0;
}

#ifdef B2T_COMMON_KERNELS   // geometry-independent kernels live in the API translation unit only
// ---------------------------------------------------------------------------------------------- exclusive scan of K
constexpr int SCAN_TPB = 1024, SCAN_ITEMS = 4, SCAN_TILE = SCAN_TPB * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* s_warp) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= d) w += t;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    const int base = warp ? s_warp[warp - 1] : 0;
    *total = s_warp[31];
    __syncthreads();
    return base + inc - v;
}
// phase 1: per-tile sums
__global__ void __launch_bounds__(SCAN_TPB) scan_tile_sums(const int32_t* K, int n, int32_t* sums) {
    __shared__ int s_warp[32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v = 0;
    for (int j = 0; j < SCAN_ITEMS; j++) if (base + j < n) v += K[base + j];
    int total;
    block_exclusive_scan(v, &total, s_warp);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}
// phase 2: scan of the tile sums by one block (tiles <= SCAN_TILE, i.e. n <= 16.7M)
__global__ void __launch_bounds__(SCAN_TPB) scan_sums(int32_t* sums, int tiles) {
    __shared__ int s_warp[32];
    const int base = threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS], t = 0;
    for (int j = 0; j < SCAN_ITEMS; j++) { v[j] = (base + j < tiles) ? sums[base + j] : 0; t += v[j]; }
    int total;
    int ex = block_exclusive_scan(t, &total, s_warp);
    for (int j = 0; j < SCAN_ITEMS; j++) { if (base + j < tiles) sums[base + j] = ex; ex += v[j]; }
}
// phase 3: offsets
// segmented argmax of the caller's per-placement scores: first maximum wins, like numpy.argmax
// rows at or past `capacity` were dropped by the expand that produced the offsets (never simulated, no score): they cannot be chosen; a
// board with no surviving row takes placement 0
__global__ void __launch_bounds__(256) select_kernel(const float* scores, const int32_t* offsets, int n, int capacity, int32_t* placement) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int a = offsets[i], b = offsets[i + 1];
    if (b > capacity) b = capacity;
    if (a > b) a = b;
    int best = 0;
    float bv = (b > a) ? scores[a] : 0.f;
    for (int r = a + 1; r < b; r++) {
        const float v = scores[r];
        if (v > bv) { bv = v; best = r - a; }
    }
    placement[i] = best;
}
// row -> {board slot, game id, placement, player} for expand: a block takes 256 consecutive boards, its threads walk the rows
// of those boards in order (coalesced 16-byte stores) and find each row's board by bisection of the offsets in shared memory
__global__ void __launch_bounds__(256) fill_row_meta(const int32_t* offsets, int n, const int32_t* idx, const uint8_t* player,
                                                     int4* row_meta, int32_t* row_board, int capacity) {
    __shared__ int s_off[257];
    __shared__ int s_gi[256];
    __shared__ uint8_t s_pl[256];
    const int b0 = blockIdx.x * 256, tid = threadIdx.x;
    const int nb = (n - b0) < 256 ? (n - b0) : 256;
    for (int t = tid; t <= nb; t += 256) s_off[t] = offsets[b0 + t];
    if (tid < nb) { s_gi[tid] = idx ? idx[b0 + tid] : b0 + tid; s_pl[tid] = player[b0 + tid]; }   // (the kernels clamp the player index)
    __syncthreads();
    const int r0 = s_off[0];
    int r1 = s_off[nb];
    if (r1 > capacity) r1 = capacity;
    for (int r = r0 + tid; r < r1; r += 256) {
        int lo = 0, hi = nb;              // s_off[lo] <= r < s_off[hi]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= r) lo = mid; else hi = mid;
        }
        row_meta[r] = make_int4(b0 + lo, s_gi[lo], r - s_off[lo], (int)s_pl[lo]);
        if (row_board) row_board[r] = b0 + lo;
    }
}
__global__ void __launch_bounds__(SCAN_TPB) scan_finish(const int32_t* K, int n, const int32_t* sums, int32_t* offsets) {
    __shared__ int s_warp[32];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS], t = 0;
    for (int j = 0; j < SCAN_ITEMS; j++) { v[j] = (base + j < n) ? K[base + j] : 0; t += v[j]; }
    int total;
    int ex = block_exclusive_scan(t, &total, s_warp) + sums[blockIdx.x];
    for (int j = 0; j < SCAN_ITEMS; j++) {
        if (base + j < n) offsets[base + j] = ex;
        ex += v[j];
        if (base + j == n - 1) offsets[n] = ex;
    }
}

// ---------------------------------------------------------------------------------------------- tables, init, misc
struct PieceMap { uint8_t m[8]; };
__global__ void __launch_bounds__(64) fill_tables_kernel(uint8_t* piece, uint32_t plen, uint8_t* hole, uint32_t hlen, PieceMap map,
                                                         int only_zs, int width) {
    const uint32_t seed = blockIdx.x * blockDim.x + threadIdx.x;
    if (seed >= (uint32_t)RNG_SEEDS) return;
    MT19937 mt;
    fill_stream_row(mt, seed, map.m, only_zs != 0, width, piece + (size_t)seed * plen, plen, hole + (size_t)seed * hlen, hlen);
}

#endif  // B2T_COMMON_KERNELS

template <class G>
__global__ void __launch_bounds__(TPB) init_kernel(uint4* state, int n_games) {
    typedef typename G::col_t col_t;
    typedef SmemLayout<G> L;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, gi = blockIdx.x * TPB + tid;
    unsigned char* wbase = smem + warp * L::WARP_BYTES;
    Game<G, WS> g;
    for (int pl = 0; pl < 2; pl++) {
        g.b[pl].cols = reinterpret_cast<col_t*>(wbase) + pl * G::NCOL * WS + lane;
        g.b[pl].stream = reinterpret_cast<uint32_t*>(wbase + L::O_STREAM) + pl * G::SW * WS + lane;
        g.p[pl].gq = reinterpret_cast<uint32_t*>(wbase + L::O_GQ) + pl * GQ_WORDS * WS + lane;
    }
    if (gi >= n_games) return;
    game_init(g);
    game_pack(g);
#pragma unroll
    for (int v = 0; v < G::NV; v++) {
        uint4 q;
        q.x = game_get_word(g, 4 * v + 0); q.y = game_get_word(g, 4 * v + 1);
        q.z = game_get_word(g, 4 * v + 2); q.w = game_get_word(g, 4 * v + 3);
        state[state_index<G>(gi, v)] = q;
    }
}

#ifdef B2T_COMMON_KERNELS
// export/import one game's NV vectors <-> contiguous buffer
__global__ void gather_game_kernel(const uint4* state, int game, int nv, uint4* out) {      // tile layout, see b2t_game.cuh
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nv) out[v] = state[((size_t)(game >> 5) * nv + v) * 32 + (game & 31)];
}
__global__ void scatter_game_kernel(uint4* state, int game, int nv, const uint4* in) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nv) state[((size_t)(game >> 5) * nv + v) * 32 + (game & 31)] = in[v];
}
// one game's colour-plane words <-> contiguous buffer (tile-of-32 layout, see KParams::tiles); copy between two games
__global__ void gather_tiles_kernel(const uint32_t* tiles, int game, int nw, uint32_t* out) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < nw) out[w] = tiles[((size_t)(game >> 5) * nw + w) * 32 + (game & 31)];
}
__global__ void scatter_tiles_kernel(uint32_t* tiles, int game, int nw, const uint32_t* in) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < nw) tiles[((size_t)(game >> 5) * nw + w) * 32 + (game & 31)] = in[w];
}
__global__ void copy_tiles_kernel(uint32_t* dst_tiles, int dst, const uint32_t* src_tiles, int src, int nw) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < nw) dst_tiles[((size_t)(dst >> 5) * nw + w) * 32 + (dst & 31)] = src_tiles ? src_tiles[((size_t)(src >> 5) * nw + w) * 32 + (src & 31)] : 0u;
}
// n games' vectors <-> one contiguous buffer [n][nv] (batched export / import: get_state of a whole vector of envs)
__global__ void gather_games_kernel(const uint4* state, int n_games, const int32_t* idx, int n, int nv, uint4* out) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * nv) return;
    const int i = (int)(t / nv), v = (int)(t % nv), game = idx ? idx[i] : i;
    out[t] = (unsigned)game < (unsigned)n_games ? state[((size_t)(game >> 5) * nv + v) * 32 + (game & 31)] : make_uint4(0u, 0u, 0u, 0u);
}
__global__ void scatter_games_kernel(uint4* state, int n_games, const int32_t* idx, int n, int nv, const uint4* in) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * nv) return;
    const int i = (int)(t / nv), v = (int)(t % nv), game = idx ? idx[i] : i;
    if ((unsigned)game < (unsigned)n_games) state[((size_t)(game >> 5) * nv + v) * 32 + (game & 31)] = in[t];
}
// PythonHandle.copy()/set() of one game, possibly between two handles of the same field size
__global__ void copy_game_kernel(uint4* dst_state, int dst, const uint4* src_state, int src, int nv) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < nv) dst_state[((size_t)(dst >> 5) * nv + v) * 32 + (dst & 31)] = src_state[((size_t)(src >> 5) * nv + v) * 32 + (src & 31)];
}

#endif  // B2T_COMMON_KERNELS

// ---------------------------------------------------------------------------------------------- any number of players
// n_players != 2 (PythonHandle takes any count, PythonHandle.cpp:5-25; no BASELINE configuration or preset of the reference uses
// one): the same rules (b2t_core.cuh: finish_actions_n and friends) run by a plain kernel — one thread per game or simulated
// placement, every player's board in shared memory, every player's scalars in local memory, byte-wise outputs.  It serves reset,
// step (keys or placement index, fused reset), observe, expand and the pokes for handles created with 1, 3 or 4 players; the
// enumerate kernel is shared.  HBM layout: the tile-of-32 layout with nv = ceil(np*PW/4) vectors per game; a player's 24th
// scalar word holds the float GamePlay::incoming_lines, the spare bits 26..29 of player 0's first scalar word the game flags.
constexpr int NTPB = 64;
template <class G> struct SmemLayoutN {
    static constexpr int PER_PLAYER = G::NCOL * (int)sizeof(typename G::col_t) + (G::SW + GQ_WORDS) * 4;   // bytes per thread
    static constexpr int total(int np) { return np * PER_PLAYER * NTPB + 128; }
};

template <class G>
__global__ void __launch_bounds__(NTPB) gamen_kernel(const KParams p) {
    typedef typename G::col_t col_t;
    constexpr int KW = G::KEYW, W0 = G::W * G::CW, SCAL0 = W0 + G::SW;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, np = p.np;
    uint16_t* s_tab = reinterpret_cast<uint16_t*>(smem);
    col_t* s_cols = reinterpret_cast<col_t*>(smem + 128);
    uint32_t* s_stream = reinterpret_cast<uint32_t*>(smem + 128 + (size_t)np * G::NCOL * NTPB * sizeof(col_t));
    uint32_t* s_gq = s_stream + (size_t)np * G::SW * NTPB;
    if (tid < 28) { s_tab[tid] = p.tabs.pcn[tid]; s_tab[32 + tid] = p.tabs.pcc[tid]; }
    Tabs tb;
    tb.pcn = s_tab; tb.pcc = s_tab + 32;
    tb.piece = p.tabs.piece; tb.hole = p.tabs.hole; tb.powtab = p.tabs.powtab;
    tb.piece_len = p.tabs.piece_len; tb.hole_len = p.tabs.hole_len; tb.spawn_rots = p.tabs.spawn_rots;
    PS ps[MAX_PLAYERS];
    Board<G, NTPB> bd[MAX_PLAYERS];
    float inc[MAX_PLAYERS];
    for (int pl = 0; pl < np; pl++) {
        bd[pl].cols = s_cols + (size_t)pl * G::NCOL * NTPB + tid;
        bd[pl].stream = s_stream + (size_t)pl * G::SW * NTPB + tid;
        ps[pl].gq = s_gq + (size_t)pl * GQ_WORDS * NTPB + tid;
        board_init_walls<G, NTPB>(bd[pl]);
        inc[pl] = 0.f;
    }
    __syncthreads();
    const uint32_t flags = p.flags;
    const bool expand = (flags & KF_EXPAND) != 0, stepping = (flags & (KF_STEP | KF_EXPAND)) != 0;
    const int row = blockIdx.x * NTPB + tid;
    int total_rows = p.n, i = row, player = 0, sel = 0, gi = 0;
    if (expand) {
        total_rows = p.offsets[p.n];
        if (total_rows > p.capacity) total_rows = p.capacity;
    }
    if (row >= total_rows) return;
    if (expand) {
        const int4 m = p.row_meta[row];
        i = m.x; gi = m.y; sel = m.z; player = m.w;
    } else gi = p.idx ? p.idx[i] : i + p.game0;
    if ((unsigned)gi >= (unsigned)p.n_games) return;       // a foreign id: an inert row
    // the game's nv vectors travel as 128-bit accesses (a warp's 32 games: 512 contiguous bytes per vector) through a per-thread
    // word array in local memory, which the per-player loops below index freely
    constexpr int LWORDS = 4 * ((MAX_PLAYERS * G::PW + 3) / 4);
    uint32_t words[LWORDS];
    uint4* gvec = p.state + ((size_t)(gi >> 5) * (size_t)p.nv) * 32 + (size_t)(gi & 31);
    auto widx = [&](int wi) { return wi; };
    int gf = 0;
    if (!(flags & KF_INIT))
        for (int v = 0; v < p.nv; v++) {
            const uint4 q = gvec[(size_t)v * 32];
            words[4 * v] = q.x; words[4 * v + 1] = q.y; words[4 * v + 2] = q.z; words[4 * v + 3] = q.w;
        }
    if (flags & KF_INIT) {
        for (int pl = 0; pl < np; pl++) {
            board_clear<G, NTPB>(bd[pl]);
            uint32_t z[14];
            for (int k = 0; k < 14; k++) z[k] = 0u;
            ps_unpack(ps[pl], z);
            for (int j = 0; j < GQ_WORDS; j++) ps[pl].gq[j * NTPB] = 0u;
            ps[pl].kind = 7; ps[pl].min_remaining = 1000; ps[pl].drop_delay = 1000;
        }
    } else {
        for (int pl = 0; pl < np; pl++) {
            const int base = pl * G::PW;
            for (int c = 0; c < G::W; c++) {
                if (G::CW == 1) bd[pl].col(c) = (col_t)words[widx(base + c)];
                else bd[pl].col(c) = (col_t)((uint64_t)words[widx(base + 2 * c)] | ((uint64_t)words[widx(base + 2 * c + 1)] << 32));
            }
            for (int j = 0; j < G::SW; j++) bd[pl].sw(j) = words[widx(base + W0 + j)];
            uint32_t sc[14];
            for (int k = 0; k < 14; k++) sc[k] = words[widx(base + SCAL0 + k)];
            if (pl == 0) { gf = (int)(sc[0] >> 26) & 15; sc[0] &= 0x03FFFFFFu; }
            ps_unpack(ps[pl], sc);
            for (int j = 0; j < GQ_WORDS; j++) ps[pl].gq[j * NTPB] = words[widx(base + SCAL0 + 14 + j)];
            inc[pl] = __uint_as_float(words[widx(base + SCAL0 + 23)]);
        }
    }
    uint32_t seed16 = ps[0].hi12 & 0xFFFFu;
    int flags_before = 0;
    for (int pl = 0; pl < np; pl++) flags_before |= ps[pl].flags;
    auto do_reset = [&](uint32_t seed) {
        seed16 = seed & 0xFFFFu;
        reset_game_n<G, NTPB>(ps, bd, inc, np, tb, seed16, gf);
        ps[0].episode = (ps[0].episode + 1u) & 0xFFFFu;
        ps[0].hi12 = seed16;
    };
    if (flags & KF_RESET) do_reset(p.seeds[i]);
    if ((flags & KF_POKE_DEAD) && p.poke_player >= 0 && p.poke_player < np) ps[p.poke_player].dead = p.poke_value;
    if ((flags & KF_CLEAR_FIELD) && p.poke_player >= 0 && p.poke_player < np) board_clear<G, NTPB>(bd[p.poke_player]);
    if (stepping) {
        if (!expand) {
            player = p.player[i];
            if (p.placement) sel = p.placement[i];
        }
        if (player >= np) player = np - 1;
        uint32_t kw[KW];
        int nkeys = 1;
        if (expand || p.placement) {
            const int kc = p.kcache[gi];
            const bool valid = sel >= 0 && sel < kc && sel < G::KMAX;   // out of range -> null action (action_list.py:27-33)
            const uint32_t* src = p.plc + ((size_t)gi * G::KMAX + (valid ? sel : 0)) * KW;
            for (int j = 0; j < KW; j++) kw[j] = valid ? src[j] : 0u;
            nkeys = packed_len<KW>(kw);
        } else {
            const bool zterm = p.key_len == nullptr;
            nkeys = zterm ? KW * 8 : (int)p.key_len[i];
            if (nkeys > KW * 8) nkeys = KW * 8;
            if (nkeys > p.lmax) nkeys = p.lmax;
            const uint8_t* kp = p.keys + (size_t)i * (size_t)p.lmax;
            for (int j = 0; j < KW; j++) {
                uint32_t w = 0;
                for (int b = 0; b < 8; b++) {
                    const int at = 8 * j + b;
                    if (at < nkeys) { const uint32_t k = kp[at]; w |= (k <= 10u ? k : 0u) << (4 * b); }
                }
                kw[j] = w;
            }
            if (zterm) { nkeys = packed_len<KW>(kw); if (nkeys > p.lmax) nkeys = p.lmax; }
        }
        const bool was_over = (gf & GF_ROUND_OVER) != 0;
        if (!was_over && !ps[player].dead) apply_keys<G, NTPB, 1>(ps[player], bd[player], tb, kw, nkeys);   // make_actions :138-147
        const bool finalize = (flags & KF_FINALIZE) != 0;
        if (finalize) {
            const bool over = finish_actions_n<G, NTPB>(ps, bd, inc, np, tb, seed16, p.elapsed, gf);
            gf = over ? (gf | GF_ENV_DONE) : (gf & ~GF_ENV_DONE);
        }
        const bool done = (gf & GF_ENV_DONE) != 0;
        const bool you_dead = np > 1 && ps[opponent_index(player, np)].dead;
        if (p.reward) p.reward[row] = (int8_t)reward_base(done, ps[player].dead, you_dead);
        if (p.done) p.done[row] = done ? 1 : 0;
        if (p.lines)
            for (int pl = 0; pl < np; pl++) p.lines[(size_t)row * np + pl] = (uint8_t)ps[pl].reward;
        if (expand) {
            if (p.key_len_out) p.key_len_out[row] = (uint8_t)nkeys;
            if (p.keys_out)
                for (int j = 0; j < p.lmax_out; j++)
                    p.keys_out[(size_t)row * p.lmax_out + j] = (j < KW * 8) ? (uint8_t)((kw[j >> 3] >> (4 * (j & 7))) & 15u) : 0;
        } else if (finalize && !was_over && p.stats) {
            unsigned long long* st = p.stats + (blockIdx.x % STAT_SLICES) * STAT_WORDS;
            atomicAdd(&st[0], 1ull);
            int lines = 0, alive = 0, last = -1;
            for (int pl = 0; pl < np; pl++) { lines += ps[pl].reward; if (!ps[pl].dead) { alive++; last = pl; } }
            if (lines) atomicAdd(&st[5], (unsigned long long)lines);
            if (done) {
                atomicAdd(&st[1], 1ull);
                if (alive == 0) atomicAdd(&st[4], 1ull);
                else if (alive == 1 && last < 2) atomicAdd(&st[2 + last], 1ull);
            }
        }
        if (!expand && done && (p.reset_seeds || p.auto_reset))
            do_reset(p.reset_seeds ? p.reset_seeds[i] : auto_seed(p.seed_base, (uint32_t)gi, ps[0].episode));
    }
    if (p.stats) {
        int flags_after = 0;
        for (int pl = 0; pl < np; pl++) flags_after |= ps[pl].flags;
        const int fresh = flags_after & ~flags_before;
        if (fresh & 1) atomicAdd(&p.stats[(blockIdx.x % STAT_SLICES) * STAT_WORDS + 6], 1ull);
        if (fresh & 2) atomicAdd(&p.stats[(blockIdx.x % STAT_SLICES) * STAT_WORDS + 7], 1ull);
    }
    if ((flags & KF_STORE) && !expand) {
        for (int pl = 0; pl < np; pl++) {
            const int base = pl * G::PW;
            for (int c = 0; c < G::W; c++) {
                const col_t v = bd[pl].col(c);
                if (G::CW == 1) words[widx(base + c)] = (uint32_t)v;
                else { words[widx(base + 2 * c)] = (uint32_t)v; words[widx(base + 2 * c + 1)] = (uint32_t)((uint64_t)v >> 32); }
            }
            for (int j = 0; j < G::SW; j++) words[widx(base + W0 + j)] = bd[pl].sw(j);
            uint32_t sc[SCALAR_WORDS];
            ps_pack(ps[pl], sc);
            if (pl == 0) sc[0] |= ((uint32_t)gf & 15u) << 26;
            for (int k = 0; k < 14; k++) words[widx(base + SCAL0 + k)] = sc[k];
            for (int j = 0; j < GQ_WORDS; j++) words[widx(base + SCAL0 + 14 + j)] = ps[pl].gq[j * NTPB];
            words[widx(base + SCAL0 + 23)] = __float_as_uint(inc[pl]);
        }
        for (int k = np * G::PW; k < 4 * p.nv; k++) words[k] = 0u;
        for (int v = 0; v < p.nv; v++) gvec[(size_t)v * 32] = make_uint4(words[4 * v], words[4 * v + 1], words[4 * v + 2], words[4 * v + 3]);
    }
    // ---- outputs, player-index order: [rows, np, H, W] / [rows, np, 19] / [rows, np] / [rows, np, 8]
    if (p.field) {
        uint8_t* f = p.field + (size_t)row * np * G::FIELD_BYTES;
        if (G::FIELD_BYTES % 4 == 0 && (reinterpret_cast<uintptr_t>(p.field) & 3) == 0) {     // four cells per store
            uint32_t* f4 = reinterpret_cast<uint32_t*>(f);
            for (int pl = 0; pl < np; pl++)
                for (int j = 0; j < G::SW; j++) {
                    const uint32_t w = bd[pl].sw(j);
                    for (int b = 0; b < 8 && j * 8 + b < G::FIELD_BYTES / 4; b++)
                        f4[pl * (G::FIELD_BYTES / 4) + j * 8 + b] = (((w >> (4 * b)) & 15u) * 0x00204081u) & 0x01010101u;
                }
        } else
            for (int pl = 0; pl < np; pl++)
                for (int j = 0; j < G::SW; j++) {
                    const uint32_t w = bd[pl].sw(j);
                    for (int b = 0; b < 32 && j * 32 + b < G::FIELD_BYTES; b++) f[pl * G::FIELD_BYTES + j * 32 + b] = (uint8_t)((w >> b) & 1u);
                }
    }
    if (p.vec)
        for (int pl = 0; pl < np; pl++) observe_vec(ps[pl], p.vec + ((size_t)row * np + pl) * VEC_LEN);
    if (p.piece_idx)
        for (int pl = 0; pl < np; pl++) p.piece_idx[(size_t)row * np + pl] = (uint8_t)ps[pl].kind;
    if (p.scalars)
        for (int pl = 0; pl < np; pl++) {
            int32_t* sv = p.scalars + ((size_t)row * np + pl) * 8;
            const PS& q = ps[pl];
            sv[0] = q.x; sv[1] = q.y; sv[2] = q.inc_count; sv[3] = q.remaining; sv[4] = q.combo_count; sv[5] = q.next; sv[6] = q.reward; sv[7] = q.dead;
        }
}

// ---------------------------------------------------------------------------------------------- traffic probe
// The HBM traffic of one fused step and nothing else: every thread reads its game's NV vectors and its key row, writes the vectors
// back and the warp writes 32 rows of observation bytes, with the step kernel's own access pattern (tiles, coalesced 128-bit
// accesses), block size and shared-memory footprint (so the same number of warps is resident), but no game logic.  Its bandwidth is
// what this read/write mix can reach on the device — the practical ceiling the step kernel is compared with next to the plain
// copy peak (bench.py: roofline.mix_peak).  Measurement aid only: b2t_debug_traffic.
template <class G>
__global__ void __launch_bounds__(TPB, B2T_GAME_MINB) traffic_kernel(uint4* state, int n, const uint8_t* keys, int lmax, const uint8_t* player,
                                                                     const uint8_t* key_len, const uint32_t* seeds, int8_t* reward, uint8_t* done,
                                                                     uint8_t* lines, uint8_t* field, uint8_t* vec, uint8_t* piece_idx) {
    constexpr int HW = G::H * G::W;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x * TPB + tid;
    const int warp_row0 = blockIdx.x * TPB + warp * 32;
    if (warp_row0 + 32 > n) return;                    // whole warps only: the probe is about bandwidth, not edges
    uint4 q[G::NV];
    const uint4* src = state + state_index<G>(row, 0);
#pragma unroll
    for (int v = 0; v < G::NV; v++) q[v] = src[v * 32];
    uint32_t x = player[row] + key_len[row] + seeds[row];
    if (lmax == 16) { const uint4 k = *reinterpret_cast<const uint4*>(keys + (size_t)row * 16); x += k.x ^ k.y ^ k.z ^ k.w; }
#pragma unroll
    for (int v = 0; v < G::NV; v++) x += q[v].x ^ q[v].w;
    reinterpret_cast<uint32_t*>(smem)[tid] = x;        // keeps the loads alive without touching the stored values
    uint4* dst = state + state_index<G>(row, 0);
#pragma unroll
    for (int v = 0; v < G::NV; v++) { q[v].y += 1u; dst[v * 32] = q[v]; }   // (modified: a store of the value just loaded would be elided)
    reward[row] = (int8_t)(x & 1); done[row] = (uint8_t)((x >> 1) & 1);
    *reinterpret_cast<uint16_t*>(lines + 2 * (size_t)row) = (uint16_t)x;
    *reinterpret_cast<uint16_t*>(piece_idx + 2 * (size_t)row) = (uint16_t)(x >> 3);
    const uint4 fill = make_uint4(x & 0x01010101u, (x >> 1) & 0x01010101u, (x >> 2) & 0x01010101u, (x >> 3) & 0x01010101u);
    uint4* of = reinterpret_cast<uint4*>(field + (size_t)warp_row0 * (2 * HW));
    for (int c = lane; c < 4 * HW; c += 32) of[c] = fill;                       // 32 rows x 2*HW bytes = 4*HW chunks of 16 bytes
    uint4* ov = reinterpret_cast<uint4*>(vec + (size_t)warp_row0 * (2 * VEC_LEN));
    for (int c = lane; c < 32 * 2 * VEC_LEN / 16; c += 32) ov[c] = fill;
}

// ---------------------------------------------------------------------------------------------- per-geometry launch table
struct GeomOps {
    int H, W, NV, KMAX, KEYW, LMAX, smem_game, smem_enum;
    cudaError_t (*prepare)();
    cudaError_t (*launch_game)(const KParams&, int rows, cudaStream_t);
    cudaError_t (*launch_enum)(const EParams&, cudaStream_t);
    cudaError_t (*launch_unpack)(const uint32_t*, const int32_t*, const int32_t*, int, int, int, int, uint8_t*, uint8_t*, cudaStream_t);
    cudaError_t (*launch_init)(uint4*, int, cudaStream_t);
    cudaError_t (*launch_gamen)(const KParams&, int rows, cudaStream_t);     // any player count (gamen_kernel)
    int PW;                       // state words per player
    cudaError_t (*launch_traffic)(uint4*, int, const uint8_t*, int, const uint8_t*, const uint8_t*, const uint32_t*, int8_t*, uint8_t*, uint8_t*,
                                  uint8_t*, uint8_t*, uint8_t*, int, cudaStream_t);
    void (*decode)(const uint32_t* words, b2t_game_raw* out, uint8_t* field, const uint32_t* planes);          // host codec, b2t_raw.cuh
    void (*encode)(uint32_t* words, const b2t_game_raw* in, const uint8_t* field, uint32_t* planes);
    int tile_words;               // colour-plane words per game: 2 * TILE_PLANES * SW
};

template <class G> struct Launch {
    typedef SmemLayout<G> L;
    static cudaError_t prepare() {
        cudaError_t e = cudaFuncSetAttribute(game_kernel<G, GM_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::GAME_TOTAL);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(game_kernel<G, GM_EXPAND>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::GAME_TOTAL);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(game_kernel<G, GM_STEP>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::GAME_TOTAL);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(enumerate_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, EnumCfg<G>::SMEM);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(gamen_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmemLayoutN<G>::total(MAX_PLAYERS));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(traffic_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::GAME_TOTAL);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(init_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::GAME_TOTAL);
    }
    static cudaError_t game(const KParams& p, int rows, cudaStream_t s) {
        if (rows <= 0) return cudaSuccess;
        const dim3 grid((rows + TPB - 1) / TPB);
        if ((p.obs_dtype != 0 && p.obs_dtype != 4) || p.key_fmt != 0 || p.auto_reset || ((p.flags & KF_STEP) && !p.placement && !p.key_len) ||
            (p.tiles && !(p.flags & KF_EXPAND))) game_kernel<G, GM_GENERIC><<<grid, TPB, L::GAME_TOTAL, s>>>(p);
        else if ((p.flags & ~KF_FINALIZE) == KF_EXPAND) game_kernel<G, GM_EXPAND><<<grid, TPB, L::GAME_TOTAL, s>>>(p);
        else if ((p.flags & ~KF_FINALIZE) == (KF_STEP | KF_STORE)) game_kernel<G, GM_STEP><<<grid, TPB, L::GAME_TOTAL, s>>>(p);
        else game_kernel<G, GM_GENERIC><<<grid, TPB, L::GAME_TOTAL, s>>>(p);
        return cudaGetLastError();
    }
    static cudaError_t enumerate(const EParams& p, cudaStream_t s) {
        if (p.n <= 0) return cudaSuccess;
        enumerate_kernel<G><<<(p.n + ENB - 1) / ENB, ETPB, EnumCfg<G>::SMEM, s>>>(p);
        return cudaGetLastError();
    }
    static cudaError_t unpack(const uint32_t* plc, const int32_t* kc, const int32_t* idx, int n_games, int n, int kmax, int lmax, uint8_t* keys,
                              uint8_t* key_len, cudaStream_t s) {
        const long long t = (long long)n * kmax;
        if (t <= 0) return cudaSuccess;
        unpack_keys_kernel<G><<<(unsigned)((t + 255) / 256), 256, 0, s>>>(plc, kc, idx, n_games, n, kmax, lmax, keys, key_len);
        return cudaGetLastError();
    }
    static cudaError_t init(uint4* state, int n, cudaStream_t s) {
        init_kernel<G><<<(n + TPB - 1) / TPB, TPB, L::GAME_TOTAL, s>>>(state, n);
        return cudaGetLastError();
    }
    static cudaError_t gamen(const KParams& p, int rows, cudaStream_t s) {
        if (rows <= 0) return cudaSuccess;
        gamen_kernel<G><<<(rows + NTPB - 1) / NTPB, NTPB, SmemLayoutN<G>::total(p.np), s>>>(p);
        return cudaGetLastError();
    }
    static cudaError_t traffic(uint4* state, int n, const uint8_t* keys, int lmax, const uint8_t* player, const uint8_t* key_len,
                               const uint32_t* seeds, int8_t* reward, uint8_t* done, uint8_t* lines, uint8_t* field, uint8_t* vec,
                               uint8_t* piece_idx, int step_like, cudaStream_t s) {
        // step_like: the step kernel's shared-memory footprint, i.e. its 16 resident warps per SM; otherwise as many as the registers allow
        traffic_kernel<G><<<(n + TPB - 1) / TPB, TPB, step_like ? L::GAME_TOTAL : TPB * 4, s>>>(state, n, keys, lmax, player, key_len, seeds, reward, done, lines,
                                                                            field, vec, piece_idx);
        return cudaGetLastError();
    }
    static const GeomOps* ops() {
        static const GeomOps o = {G::H, G::W, G::NV, G::KMAX, G::KEYW, G::LMAX, L::GAME_TOTAL, EnumCfg<G>::SMEM,
                                  &prepare, &game, &enumerate, &unpack, &init, &gamen, G::PW, &traffic, &decode_game<G>, &encode_game<G>, 2 * TILE_PLANES * G::SW};
        return &o;
    }
};

}  // namespace b2t
