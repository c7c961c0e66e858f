// b200tetris — game core: bit-board Tetris rules shared by every kernel.
//
// One THREAD owns one game (step) or one simulated placement (expand).  A player's board lives in shared
// memory as
//   * W column masks (bit (row+YB) of column c = cell (row,c) occupied; bits below YB = ceiling, bits above
//     the field = floor, XL/XR all-ones wall columns either side) — collision, hard drop and the "fits at
//     every height at once" mask of the placement enumerator are shifts and ANDs on these, and
//   * a row-major bit stream (bit row*W+col) that exists only so the uint8 observation tensors can be
//     expanded 16 cells per 128-bit store.
// Scalars (timers, garbage queue, combo, RNG counters) live in registers (struct PS).
//
// Everything here is __host__ __device__ so that tests/hostsim can run the identical code on the CPU against
// the oracle; the product library only ever calls it from kernels (b2t_kernels.cu).
//
// Reference semantics restated (file:line under /root/reference/environment/game_backend/source):
//   gameField.cpp:10-145, gamePlay.cpp:38-230, PythonHandle.cpp:49-188, Garbage.cpp, Combo.cpp, DropDelay.cpp,
//   randomizer.cpp, TestField.cpp:3-415 — cited again at each function.
#pragma once
#include <stdint.h>
#if defined(B2T_BOUNDS_CHECK)
#include <cstdio>
#include <cstdlib>
#endif

#if defined(__CUDACC__)
#define B2T_HD __host__ __device__ __forceinline__
#else
#define B2T_HD inline
#endif

namespace b2t {

// ------------------------------------------------------------------------------------------------ bit helpers
B2T_HD int ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)v) - 1;
#else
    return __builtin_ctz(v);
#endif
}
B2T_HD int ctz64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
B2T_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
B2T_HD int popc64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}
B2T_HD int msb32(uint32_t v) {  // index of highest set bit, v != 0
#if defined(__CUDA_ARCH__)
    return 31 - __clz((int)v);
#else
    return 31 - __builtin_clz(v);
#endif
}
B2T_HD int msb64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return 63 - __clzll((long long)v);
#else
    return 63 - __builtin_clzll(v);
#endif
}
B2T_HD int ctz(uint32_t v) { return ctz32(v); }
B2T_HD int ctz(uint64_t v) { return ctz64(v); }
B2T_HD int popc(uint32_t v) { return popc32(v); }
B2T_HD int popc(uint64_t v) { return popc64(v); }
B2T_HD int msb(uint32_t v) { return msb32(v); }
B2T_HD int msb(uint64_t v) { return msb64(v); }

// exact IEEE single ops without FMA contraction (Combo.cpp:40)
B2T_HD float f_add(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b; return r;
#endif
}
B2T_HD float f_mul(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b; return r;
#endif
}
B2T_HD float f_div(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    volatile float r = a / b; return r;
#endif
}
B2T_HD double d_mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b; return r;
#endif
}
B2T_HD int32_t d_to_i32(double v) {
#if defined(__CUDA_ARCH__)
    return __double2int_rz(v);
#else
    return (int32_t)v;
#endif
}

template <bool C, class A, class B> struct pick_t { typedef A type; };
template <class A, class B> struct pick_t<false, A, B> { typedef B type; };

// ------------------------------------------------------------------------------------------------ geometry
constexpr int GQ_CAP = 6;        // garbage-queue capacity (entries); overflow raises a sticky per-game flag
constexpr int SCALAR_WORDS = 24; // per player
constexpr int VEC_LEN = 19;      // x,y,incoming_lines,combo_time,combo_count,nextpiece[7],piece[7]
constexpr int RNG_SEEDS = 65536; // the reference truncates its seed to a short (randomizer.cpp:34-36,47-49)

template <int H_, int W_> struct Geo {
    static_assert(W_ >= 4 && W_ < 32 && H_ >= 4 && H_ + 3 + 4 <= 64, "unsupported field size");
    static constexpr int H = H_, W = W_;
    static constexpr int YB = 3;                    // row r lives at bit r+YB; rows -3..-1 are ceiling
    static constexpr bool WIDE = (H_ + YB + 4 > 32); // 4-row piece at y=H must still fit in the word
    typedef typename pick_t<WIDE, uint64_t, uint32_t>::type col_t;
    static constexpr int CBITS = WIDE ? 64 : 32;
    static constexpr int CW = WIDE ? 2 : 1;         // 32-bit words per column in the stored state
    static constexpr int XL = 4, XR = 5;            // wall columns: x in [-4, W+1], +3 piece columns
    static constexpr int NCOL = W_ + XL + XR;
    static constexpr int SW = (H_ * W_ + 31) / 32;  // stream words per player
    static constexpr int PW = W_ * CW + SW + SCALAR_WORDS; // state words per player
    static constexpr int GWORDS = 2 * PW;
    static constexpr int NV = (GWORDS + 3) / 4;     // uint4 vectors per game in HBM
    static constexpr int SPAWN_X = (W_ - 4) / 2;
    static constexpr int FIELD_BYTES = H_ * W_;
    static constexpr int OBS_FIELD = 2 * H_ * W_;   // bytes of field observation per game
    // A key sequence is at most: rotation + (W-4)/2+1 lateral + [turn] + one lateral per column crossed (<= W-1) with the rows
    // risen in between (their sum <= H) + the final 7  <=  H + 2W + 4 tokens.  4-bit tokens, whole uint4s.
    static constexpr int KEYW = (((H_ + 2 * W_ + 4 + 7) / 8 + 3) / 4) * 4;
    static constexpr int LMAX = KEYW * 8;
    static constexpr int KMAX = (W_ > 10) ? 256 : 96;
    B2T_HD static col_t one() { return (col_t)1; }
    B2T_HD static col_t ceil_mask() { return (col_t)((1u << YB) - 1u); }
    B2T_HD static col_t rows_mask() { return (col_t)(((one() << H) - 1) << YB); }
    B2T_HD static col_t floor_mask() { return (col_t)~(((one() << (H + YB)) - 1)); }
    B2T_HD static col_t empty_col() { return (col_t)(ceil_mask() | floor_mask()); }
    B2T_HD static col_t wall_col() { return (col_t)~(col_t)0; }
};

// ------------------------------------------------------------------------------------------------ piece tables
// 7 pieces x 4 rotations, index kind*4+rot.  Built once on the host by make_piece_tables() from the rotation-0
// cell lists and the rotation rule of pieces.cpp:5-53 (I and O turn as 4x4, the rest as the top-left 3x3).
struct PieceTables {
    uint16_t pcn[28];  // nibble c = rows occupied in grid column c
    uint16_t pcc[28];  // 4 cells, nibble i = cx | cy<<2
    uint8_t spawn_rot[8];
};

inline void make_piece_tables(PieceTables& t) {
    static const uint8_t base[7][4][2] = {  // (row, col) at current_rotation 0, gamePlay.cpp:125-137
        {{0, 1}, {1, 1}, {2, 1}, {2, 2}}, {{0, 1}, {1, 1}, {2, 0}, {2, 1}}, {{0, 1}, {1, 1}, {1, 2}, {2, 2}},
        {{0, 1}, {1, 0}, {1, 1}, {2, 0}}, {{0, 1}, {1, 1}, {2, 1}, {3, 1}}, {{1, 0}, {1, 1}, {1, 2}, {2, 1}},
        {{1, 1}, {1, 2}, {2, 1}, {2, 2}}};
    static const uint8_t spawn[7] = {3, 1, 3, 1, 1, 2, 0};  // gamePlay.cpp:117
    for (int k = 0; k < 7; k++) {
        int n = (k == 4 || k == 6) ? 4 : 3;
        int cell[4][2];
        for (int i = 0; i < 4; i++) { cell[i][0] = base[k][i][0]; cell[i][1] = base[k][i][1]; }
        for (int r = 0; r < 4; r++) {
            uint16_t pcn = 0, pcc = 0;
            // canonical cell order: column-major scan like BasicField::possible (gameField.cpp:11-13)
            int ord = 0;
            for (int cx = 0; cx < 4; cx++)
                for (int cy = 0; cy < 4; cy++)
                    for (int i = 0; i < 4; i++)
                        if (cell[i][0] == cy && cell[i][1] == cx) {
                            pcn |= (uint16_t)(1u << (4 * cx + cy));
                            pcc |= (uint16_t)((cx | (cy << 2)) << (4 * ord++));
                        }
            t.pcn[k * 4 + r] = pcn;
            t.pcc[k * 4 + r] = pcc;
            for (int i = 0; i < 4; i++) {  // rcw: new[x][n-1-y] = old[y][x]
                int y = cell[i][0], x = cell[i][1];
                cell[i][0] = x;
                cell[i][1] = n - 1 - y;
            }
        }
        t.spawn_rot[k] = spawn[k];
    }
    t.spawn_rot[7] = 0;
}

// Read-only tables every kernel needs.  pcn/pcc point at a shared-memory copy on the device.
struct Tabs {
    const uint16_t* pcn;
    const uint16_t* pcc;
    const uint8_t* piece;  // [RNG_SEEDS][piece_len]  mapped piece stream (gamePlay.cpp:218-230, randomizer.cpp:10-32)
    const uint8_t* hole;   // [RNG_SEEDS][hole_len]   garbage hole stream (randomizer.cpp:38-45)
    const double* powtab;  // [256] pow(cc, 1.4+0.01cc) tabulated on the host with libm (Combo.cpp:41)
    uint32_t piece_len, hole_len;
    uint32_t spawn_rots;   // 7 x 2 bits
};
B2T_HD int spawn_rot_of(const Tabs& t, int kind) { return (int)((t.spawn_rots >> (2 * kind)) & 3u); }

// ------------------------------------------------------------------------------------------------ board view
template <class G, int S> struct Board {
    typedef typename G::col_t col_t;
    col_t* cols;        // column c at cols[(c + XL) * S]
    uint32_t* stream;   // word w at stream[w * S]
    // Optional tile-colour planes (b2t_config.flags & B2T_FLAG_TILES): the raw cell values 1..8 of PythonHandle.states[p].field
    // (piece.tile = piece + 1, gameField.cpp:105-110; garbage = 8, gamePlay.cpp:202) as three bit streams of (value - 1) in the
    // stream's bit order, word w of plane k at tiles[(k*SW + w) * S].  nullptr (the default) = occupancy only.
    uint32_t* tiles = nullptr;
    B2T_HD col_t& col(int c) const { return cols[(c + G::XL) * S]; }
    B2T_HD uint32_t& sw(int w) const { return stream[w * S]; }
    B2T_HD uint32_t& tw(int k, int w) const { return tiles[(k * G::SW + w) * S]; }
};
constexpr int TILE_PLANES = 3;
// accessors so that one routine shifts the occupancy stream and every colour plane alike
template <class G, int S> struct StreamRef {
    uint32_t* p;
    B2T_HD uint32_t& operator()(int w) const { return p[w * S]; }
};

template <class G, int S> B2T_HD void board_init_walls(const Board<G, S>& b) {
    for (int c = -G::XL; c < 0; c++) b.col(c) = G::wall_col();
    for (int c = G::W; c < G::W + G::XR; c++) b.col(c) = G::wall_col();
}
template <class G, int S> B2T_HD void board_clear(const Board<G, S>& b) {
    for (int c = 0; c < G::W; c++) b.col(c) = G::empty_col();
    for (int w = 0; w < G::SW; w++) b.sw(w) = 0;
    if (b.tiles)
        for (int w = 0; w < TILE_PLANES * G::SW; w++) b.tiles[w * S] = 0;
}

// gameField.cpp:10-20 possible(): pcn = column nibbles of the piece at its rotation
#if defined(B2T_BOUNDS_CHECK) && !defined(__CUDA_ARCH__)
#define B2T_CHECK_XY(x, y) do { if ((x) < -G::XL || (x) + 3 > G::W + G::XR - 1 || (y) + G::YB < 0 || (y) + G::YB + 4 > G::CBITS) { \
    fprintf(stderr, "b2t bounds: x=%d y=%d (W=%d H=%d) at %s:%d\n", (int)(x), (int)(y), G::W, G::H, __FILE__, __LINE__); abort(); } } while (0)
#else
#define B2T_CHECK_XY(x, y) do { } while (0)
#endif

template <class G, int S> B2T_HD bool fits(const Board<G, S>& b, int x, int y, uint32_t pcn) {
    typedef typename G::col_t col_t;
    B2T_CHECK_XY(x, y);
    const col_t* p = b.cols + (x + G::XL) * S;
    const int sh = y + G::YB;
    col_t acc = (p[0] >> sh) & (col_t)(pcn & 15u);
    acc |= (p[S] >> sh) & (col_t)((pcn >> 4) & 15u);
    acc |= (p[2 * S] >> sh) & (col_t)((pcn >> 8) & 15u);
    acc |= (p[3 * S] >> sh) & (col_t)((pcn >> 12) & 15u);
    return acc == 0;
}
// bit (y+YB) set  <=>  the piece collides at height y (all heights at once); pcc = cell list
template <class G, int S> B2T_HD typename G::col_t collide_mask(const Board<G, S>& b, int x, uint32_t pcc) {
    typedef typename G::col_t col_t;
    B2T_CHECK_XY(x, 0);
    const col_t* p = b.cols + (x + G::XL) * S;
    col_t m = p[(pcc & 3u) * S] >> ((pcc >> 2) & 3u);
    m |= p[((pcc >> 4) & 3u) * S] >> ((pcc >> 6) & 3u);
    m |= p[((pcc >> 8) & 3u) * S] >> ((pcc >> 10) & 3u);
    m |= p[((pcc >> 12) & 3u) * S] >> ((pcc >> 14) & 3u);
    return m;
}
// gameField.cpp:49-53 hd(): lowest y' >= y reachable by moving down until the next step collides
template <class G> B2T_HD int rest_from(typename G::col_t cm, int y) { return y + ctz((typename G::col_t)(cm >> (y + G::YB + 1))); }

// gameField.cpp:105-110 addPiece (occupancy only) — columns and stream
template <class G, int S> B2T_HD void stamp(const Board<G, S>& b, int x, int y, uint32_t pcn, uint32_t pcc, int kind) {
    typedef typename G::col_t col_t;
    col_t* p = b.cols + (x + G::XL) * S;
    const int sh = y + G::YB;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        uint32_t nib = (pcn >> (4 * c)) & 15u;
        if (nib) p[c * S] |= (col_t)nib << sh;   // nib==0 columns may be walls: leave them alone
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int cx = (pcc >> (4 * i)) & 3, cy = (pcc >> (4 * i + 2)) & 3;
        int bit = (y + cy) * G::W + x + cx;
        b.sw(bit >> 5) |= 1u << (bit & 31);
        if (b.tiles)                                   // cell value = piece.tile = kind + 1: the planes hold `kind`; a piece that
            for (int k = 0; k < TILE_PLANES; k++) {    // dies on spawn overwrites the cells it overlaps (gameField.cpp:105-110)
                uint32_t& t = b.tw(k, bit >> 5);
                t = (t & ~(1u << (bit & 31))) | ((uint32_t)((kind >> k) & 1) << (bit & 31));
            }
    }
}

// row r of a row-major bit stream disappears: bits [0, r*W) move up by W, bits [(r+1)*W, ..) stay
template <class G, class A> B2T_HD void stream_remove_row(A sw, int r) {
    const int cut = r * G::W, end = cut + G::W;
    uint32_t prev = 0;  // bits shifted out of the previous (lower) word
    for (int w = 0; w < G::SW; w++) {
        const int lo = w * 32;
        uint32_t v = sw(w);
        // part of this word below `cut`
        uint32_t mlow = cut <= lo ? 0u : (cut >= lo + 32 ? 0xFFFFFFFFu : ((1u << (cut - lo)) - 1u));
        uint32_t mkeep = end <= lo ? 0xFFFFFFFFu : (end >= lo + 32 ? 0u : ~((1u << (end - lo)) - 1u));
        uint32_t lowpart = v & mlow;
        uint32_t shifted;
        if (G::W < 32) shifted = (lowpart << G::W) | prev;
        else shifted = prev;
        uint32_t nprev = (G::W < 32) ? (lowpart >> (32 - G::W)) : lowpart;
        // shifted bits may only land below `end`
        uint32_t mdst = end <= lo ? 0u : (end >= lo + 32 ? 0xFFFFFFFFu : ((1u << (end - lo)) - 1u));
        sw(w) = (v & mkeep) | (shifted & mdst);
        prev = nprev;
    }
}
// every row moves up by one (the top row is lost) and `row` becomes the bottom row
template <class G, class A> B2T_HD void stream_push_row(A sw, uint32_t row) {
    for (int w = 0; w < G::SW; w++) {
        uint32_t v = sw(w), nx = (w + 1 < G::SW) ? sw(w + 1) : 0u;
        sw(w) = (G::W < 32) ? ((v >> G::W) | (nx << (32 - G::W))) : nx;
    }
    const int bit = (G::H - 1) * G::W;
    const int total = G::H * G::W;
    for (int w = bit >> 5; w < G::SW; w++) {
        const int lo = w * 32;
        uint32_t ins = (lo >= bit) ? (row >> (lo - bit)) : (row << (bit - lo));
        uint32_t m = 0xFFFFFFFFu;
        if (lo < bit) m &= ~((1u << (bit - lo)) - 1u);
        if (total < lo + 32) m &= (1u << (total - lo)) - 1u;
        sw(w) = (sw(w) & ~m) | (ins & m);
    }
}
template <class G, class A> B2T_HD uint32_t stream_bottom_row(A sw) {   // the W bits of row H-1
    const int bit = (G::H - 1) * G::W;
    uint32_t row = 0;
    for (int w = bit >> 5; w < G::SW; w++) {
        const int lo = w * 32;
        row |= (lo >= bit) ? (sw(w) << (lo - bit)) : (sw(w) >> (bit - lo));
    }
    return row & ((G::W < 32) ? ((1u << G::W) - 1u) : 0xFFFFFFFFu);
}

// remove one row: everything above slides down, row 0 becomes empty (gameField.cpp:112-118)
template <class G, int S> B2T_HD void remove_row(const Board<G, S>& b, int r) {
    typedef typename G::col_t col_t;
    const col_t low = (col_t)(((G::one() << (r + G::YB)) - 1) & ~G::ceil_mask());
    const col_t keep = (col_t)~((G::one() << (r + G::YB + 1)) - 1);
    for (int c = 0; c < G::W; c++) {
        col_t v = b.col(c);
        b.col(c) = (col_t)((v & keep) | ((v & low) << 1) | G::ceil_mask());
    }
    stream_remove_row<G>(StreamRef<G, S>{b.stream}, r);
    if (b.tiles)
        for (int k = 0; k < TILE_PLANES; k++) stream_remove_row<G>(StreamRef<G, S>{b.tiles + k * G::SW * S}, r);
}

// gameField.cpp:120-145 clearlines(): rows posY..H-1, bottom-up with re-check  ==  repeatedly remove the lowest full row
template <class G, int S> B2T_HD int clear_lines(const Board<G, S>& b, int pos_y) {
    typedef typename G::col_t col_t;
    int lo = pos_y < 0 ? 0 : pos_y;
    if (lo > G::H) lo = G::H;
    const col_t range = (col_t)(G::rows_mask() & ~((G::one() << (lo + G::YB)) - 1));
    int lines = 0;
    for (;;) {
        col_t full = range;
        for (int c = 0; c < G::W; c++) full &= b.col(c);
        if (!full) break;
        remove_row<G, S>(b, msb(full) - G::YB);
        lines++;
    }
    return lines;
}

// gamePlay.cpp:199-204 addGarbageLine(hole): see SURVEY A.7-4 for the `x<10` quirk
template <class G, int S> B2T_HD void garbage_row(const Board<G, S>& b, int hole) {
    typedef typename G::col_t col_t;
    const col_t bottom = G::one() << (G::H - 1 + G::YB);
    uint32_t old_row = 0;
    for (int c = 0; c < G::W; c++) {
        col_t v = b.col(c);
        const bool was = (v & bottom) != 0;
        if (was) old_row |= 1u << c;
        v = (col_t)((v >> 1) | G::floor_mask() | G::ceil_mask());  // new bottom cell = 1 (from the floor)
        if (c >= 10 && !was) v &= ~bottom;
        if (c == hole) v &= ~bottom;
        b.col(c) = v;
    }
    uint32_t row = old_row;
    row |= (G::W >= 10) ? 0x3FFu : ((1u << G::W) - 1u);
    if (hole < G::W) row &= ~(1u << hole);
    if (b.tiles) {   // rows move up keeping their colours; columns 0..9 of the new bottom row become 8 (planes: 7), the rest keep the old bottom row
        const uint32_t grey = ((G::W >= 10) ? 0x3FFu : ((1u << G::W) - 1u)) & ((hole < G::W) ? ~(1u << hole) : 0xFFFFFFFFu);
        for (int k = 0; k < TILE_PLANES; k++) {
            StreamRef<G, S> pl{b.tiles + k * G::SW * S};
            uint32_t keep = stream_bottom_row<G>(pl) & ~0x3FFu;
            if (hole < G::W) keep &= ~(1u << hole);
            stream_push_row<G>(pl, keep | grey);
        }
    }
    stream_push_row<G>(StreamRef<G, S>{b.stream}, row);
}

// ------------------------------------------------------------------------------------------------ player scalars
// A player's scalars IN REGISTERS ARE the 14 state words of the HBM layout: C bit-fields give every small field its name
// while two players cost 28 registers instead of 54 and nothing is packed or unpacked.  (x86-64 and NVCC allocate bit-fields
// from the least significant bit of 32-bit units, in declaration order.)
struct PS {
    // word 0
    uint32_t kind : 3, rot : 2, next : 3;
    int32_t x : 8, y : 8;
    uint32_t dead : 1, lockdown : 1, : 6;
    // word 1
    uint32_t reward : 8, inc_count : 8, combo_count : 8, line_count : 8;
    // word 2
    uint32_t remaining : 16, pending : 16;
    // words 3..5
    int32_t time_ms, combo_start, combo_time;
    // word 6
    uint32_t drop_delay : 16, incoming : 16;
    // words 7..11
    int32_t drop_delay_time, inc_dd_time, lockdown_time, min_remaining;
    uint32_t piece_ctr;
    // word 12   (hi12: player 0 = seed16, player 1 = game flags)
    uint32_t hole_ctr : 16, hi12 : 16;
    // word 13   (episode: player 0 only — rounds this game has started, the counter behind engine-chosen reset seeds)
    uint32_t gq_n : 8, flags : 8, episode : 16;
    // GarbageHandler::garbage (Garbage.h:27): the queue is almost always empty, so it lives in (shared) memory, not in
    // registers: 9 words, word j at gq[j*S]: 0..2 = six int16 counts, 3..8 = six int32 release times (state words 14..22)
    uint32_t* gq;
};
static_assert(sizeof(PS) >= 14 * 4, "PS must start with the 14 scalar state words");
// The piece being moved by a key string, unpacked (the key loop touches these constantly)
struct Cur {
    int kind, rot, x, y, lockdown;
    int32_t time_ms, drop_delay_time, lockdown_time;
};
constexpr int GQ_WORDS = 3 + GQ_CAP;   // state words 14..22 of a player
// game flags (player 1, word 12 high half)
constexpr int GF_ROUND_OVER = 1, GF_ENV_DONE = 2, GF_WINNER_SHIFT = 2, GF_GQ_OVERFLOW = 16, GF_RNG_OVERFLOW = 32;

union PSWords {   // word view of the leading 14 words of PS
    PS s;
    uint32_t w[14];
    B2T_HD PSWords() {}
};
B2T_HD void ps_unpack(PS& s, const uint32_t* w) {
    PSWords u;
#pragma unroll
    for (int i = 0; i < 14; i++) u.w[i] = w[i];
    uint32_t* q = s.gq;
    s = u.s;
    s.gq = q;
}
B2T_HD void ps_pack(const PS& s, uint32_t* w) {
    PSWords u;
    u.s = s;
#pragma unroll
    for (int i = 0; i < 14; i++) w[i] = u.w[i];
    w[0] &= 0x03FFFFFFu;                        // padding bits are stored as zero
    w[23] = 0;   // words 14..22 (the garbage queue) never pass through registers
}

// ---- DropDelay.cpp
B2T_HD bool dd_check(PS& s, int32_t t) {  // :3-21 (increaseDropDelay is the constant 3000 of :29)
    if (t - s.inc_dd_time > 3000) {
        if (s.drop_delay > 200) s.drop_delay -= 10;
        else if (s.drop_delay > 100) s.drop_delay -= 5;
        else if (s.drop_delay > 50) s.drop_delay -= 2;
        else if (s.drop_delay > 10) s.drop_delay -= 1;
        s.inc_dd_time = t;
    }
    if (t - s.drop_delay_time > s.drop_delay) { s.drop_delay_time = t; return true; }
    return false;
}
template <class ST> B2T_HD void dd_reset(ST& s, int32_t t) { s.drop_delay_time = t; s.lockdown = 0; }                  // :23-26
template <class ST> B2T_HD void dd_set(ST& s, int32_t t) { if (!s.lockdown) s.lockdown_time = t + 400; s.lockdown = 1; }  // :37-41

// ---- Garbage.cpp (queue front is index 0); S = stride of the queue words
template <int S> B2T_HD int gq_count_at(const PS& s, int i) { return (int)(int16_t)(s.gq[(i >> 1) * S] >> (16 * (i & 1))); }
template <int S> B2T_HD void gq_set_count(PS& s, int i, int v) {
    uint32_t& w = s.gq[(i >> 1) * S];
    const int sh = 16 * (i & 1);
    w = (w & ~(0xFFFFu << sh)) | ((uint32_t)(v & 0xFFFF) << sh);
}
template <int S> B2T_HD int32_t gq_delay_at(const PS& s, int i) { return (int32_t)s.gq[(3 + i) * S]; }
template <int S> B2T_HD void gq_set_delay(PS& s, int i, int32_t v) { s.gq[(3 + i) * S] = (uint32_t)v; }
template <int S> B2T_HD void gq_pop(PS& s) {
    const uint32_t c0 = s.gq[0], c1 = s.gq[S], c2 = s.gq[2 * S];
    s.gq[0] = (c0 >> 16) | (c1 << 16); s.gq[S] = (c1 >> 16) | (c2 << 16); s.gq[2 * S] = c2 >> 16;
    for (int i = 0; i + 1 < GQ_CAP; i++) s.gq[(3 + i) * S] = s.gq[(4 + i) * S];
    s.gq[(3 + GQ_CAP - 1) * S] = 0u;
    s.gq_n--;
}
template <int S> B2T_HD void gq_clear(PS& s) {
    for (int j = 0; j < GQ_WORDS; j++) s.gq[j * S] = 0u;
    s.gq_n = 0;
}
template <int S> B2T_HD int gq_total(const PS& s) {  // :9-14
    if (s.gq_n == 0) return 0;
    int t = 0;
    for (int i = 0; i < s.gq_n; i++) t += gq_count_at<S>(s, i);
    return t & 0xFFFF;
}
template <int S> B2T_HD void gq_add(PS& s, int amount, int32_t t) {  // :22-24
    if (s.gq_n >= GQ_CAP) { s.flags |= 1; return; }  // sticky overflow, surfaced through b2t_stats
    gq_set_count<S>(s, s.gq_n, (int16_t)amount);
    gq_set_delay<S>(s, s.gq_n, t + 1000);
    s.gq_n++;
}
template <int S> B2T_HD int gq_block(PS& s, int amount, int32_t t, bool freeze) {  // :26-52
    amount &= 0xFFFF;
    if (s.gq_n == 0) return amount;
    const int32_t delay = gq_delay_at<S>(s, 0);
    while (amount && s.gq_n) {
        const int c = gq_count_at<S>(s, 0) - 1;
        gq_set_count<S>(s, 0, c);
        amount--;
        if (c == 0) gq_pop<S>(s);
    }
    if (s.gq_n) {
        int32_t d = gq_delay_at<S>(s, 0);
        if (delay > d) d = delay;
        if (freeze) {
            const int32_t a = d + 450, b = t + s.min_remaining + 450;
            d = a < b ? a : b;
        }
        gq_set_delay<S>(s, 0, d);
    } else s.min_remaining = 1000;
    return amount;
}
template <int S> B2T_HD bool gq_check(PS& s, int32_t t) {  // :54-72 (addDelay is the constant 450 of :7)
    if (s.gq_n == 0) return false;
    const int32_t front = gq_delay_at<S>(s, 0);
    if (t > front) {
        const int32_t delay = front + 450;
        const int c = gq_count_at<S>(s, 0) - 1;
        gq_set_count<S>(s, 0, c);
        if (c == 0) gq_pop<S>(s);
        if (s.gq_n) {
            int32_t d = gq_delay_at<S>(s, 0);
            if (delay > d) { d = delay; gq_set_delay<S>(s, 0, d); }
            s.min_remaining = d - t;
        } else s.min_remaining = 1000;
        return true;
    }
    const int32_t rem = front - t;
    if (rem < s.min_remaining) s.min_remaining = rem;
    return false;
}

// ---- Combo.cpp
B2T_HD void cb_increase(PS& s, int32_t t, int amount) {  // :15-30
    if (s.combo_count == 0) { s.combo_start = t; s.combo_time = 0; }
    s.combo_count = (s.combo_count + 1) & 255;
    float line_time = 0.f;
    for (int i = 0; i < amount; i++) {
        s.line_count = (s.line_count + 1) & 255;
        line_time = f_add(line_time, (float)(1000 / (s.line_count ? s.line_count : 1)));
    }
    const float add = f_add((float)(800 / (s.combo_count ? s.combo_count : 1)), line_time);
    s.combo_time = (int32_t)f_add((float)s.combo_time, add);
}
B2T_HD int cb_check(PS& s, int32_t t, const Tabs& tb) {  // :32-48
    const int32_t left = s.combo_start + s.combo_time - t;
    s.remaining = left < 0 ? 0 : (left & 0xFFFF);
    if (t > s.combo_start + s.combo_time && s.combo_count != 0) {
        const float dur = f_add(1.f, f_mul(f_div((float)t, 60000.f), 0.1f));
        const int sent = d_to_i32(d_mul(tb.powtab[s.combo_count], (double)dur)) & 0xFFFF;
        s.combo_count = 0;
        s.line_count = 0;
        return sent;
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ GamePlay
template <class G, int S> struct Side {  // one player: scalars + board view
    PS& s;
    Board<G, S> b;
};

B2T_HD uint32_t pcn_of(const Tabs& tb, int kind, int rot) { return tb.pcn[(kind & 7) * 4 + rot]; }
B2T_HD uint32_t pcc_of(const Tabs& tb, int kind, int rot) { return tb.pcc[(kind & 7) * 4 + rot]; }

// gamePlay.cpp:71-88 makeNewPiece/copyPiece; the draw is a table lookup (see b2t_rng.cuh)
template <class G, int S> B2T_HD bool make_new_piece(PS& s, const Board<G, S>& b, const Tabs& tb, uint32_t seed16) {
    s.kind = s.next;
    s.rot = spawn_rot_of(tb, s.kind);
    s.x = G::SPAWN_X;
    s.y = 0;
    uint32_t at = s.piece_ctr;
    if (at >= tb.piece_len) { at = tb.piece_len - 1; s.flags |= 2; }  // sticky: stream table exhausted
    s.next = tb.piece[(size_t)seed16 * tb.piece_len + at];
    s.piece_ctr++;
    const uint32_t pcn = pcn_of(tb, s.kind, s.rot);
    if (!fits<G, S>(b, s.x, s.y, pcn)) { stamp<G, S>(b, s.x, s.y, pcn, pcc_of(tb, s.kind, s.rot), (int)s.kind); return true; }
    return false;
}
// gamePlay.cpp:160-171 sendLines
template <int S> B2T_HD int send_lines(PS& s, int lines) {
    s.pending = (s.pending + lines) & 0xFFFF;
    if (lines == 0) { s.combo_time -= 200; return 0; }  // Combo.cpp:50-52
    const int amount = gq_block<S>(s, lines - 1, s.time_ms, true);
    cb_increase(s, s.time_ms, lines);
    return amount;
}
// gamePlay.cpp:48-52 hd_make
template <class G, int S, class ST> B2T_HD void hd_make(ST& s, const Board<G, S>& b, const Tabs& tb) {
    if (s.kind > 6) return;
    const uint32_t pcc = pcc_of(tb, s.kind, s.rot);
    s.y = rest_from<G>(collide_mask<G, S>(b, s.x, pcc), s.y);
    stamp<G, S>(b, s.x, s.y, pcn_of(tb, s.kind, s.rot), pcc, (int)s.kind);
    dd_reset(s, s.time_ms);
}
// gamePlay.cpp:54-59 hd_finish: -1 = died
template <class G, int S> B2T_HD int hd_finish(PS& s, const Board<G, S>& b, const Tabs& tb, uint32_t seed16) {
    const int sent = send_lines<S>(s, clear_lines<G, S>(b, s.y));
    if (make_new_piece<G, S>(s, b, tb, seed16)) return -1;
    return sent;
}
// gamePlay.cpp:61-69 mDown
template <class G, int S> B2T_HD bool gp_move_down(PS& s, const Board<G, S>& b, const Tabs& tb) {
    if (fits<G, S>(b, s.x, s.y + 1, pcn_of(tb, s.kind, s.rot))) { s.y++; dd_reset(s, s.time_ms); return true; }
    dd_set(s, s.time_ms);
    return false;
}
// gamePlay.cpp:179-192 pushGarbage
template <class G, int S> B2T_HD bool push_garbage(PS& s, const Board<G, S>& b, const Tabs& tb, uint32_t seed16) {
    uint32_t at = (uint32_t)s.hole_ctr;
    if (at >= tb.hole_len) { at = tb.hole_len - 1; s.flags |= 2; }
    const int hole = tb.hole[(size_t)seed16 * tb.hole_len + at];
    s.hole_ctr = (s.hole_ctr + 1) & 0xFFFF;
    garbage_row<G, S>(b, hole);
    const uint32_t pcn = pcn_of(tb, s.kind, s.rot);
    if (s.y > 0) s.y--;
    if (!fits<G, S>(b, s.x, s.y, pcn)) {
        if (s.y > 0) s.y--;
        else return true;
    }
    return false;
}
// gamePlay.cpp:90-114 delayCheck: -1 = died, else lines sent
template <class G, int S> B2T_HD int delay_check(PS& s, const Board<G, S>& b, const Tabs& tb, uint32_t seed16, int elapsed) {
    s.time_ms += elapsed;
    if (dd_check(s, s.time_ms)) gp_move_down<G, S>(s, b, tb);
    if (s.lockdown && s.time_ms > s.lockdown_time && !gp_move_down<G, S>(s, b, tb)) {  // DropDelay.cpp:43-48
        hd_make<G, S>(s, b, tb);                                                        // gamePlay.cpp:38-46 hd()
        return hd_finish<G, S>(s, b, tb, seed16);
    }
    if (s.incoming >= 1) { gq_add<S>(s, s.incoming, s.time_ms); s.incoming = 0; }
    int sent = cb_check(s, s.time_ms, tb);
    if (sent) sent = gq_block<S>(s, sent, s.time_ms, false);
    if (gq_check<S>(s, s.time_ms))
        if (push_garbage<G, S>(s, b, tb, seed16)) return -1;
    return sent;
}

// ---- keys (PythonHandle.cpp:73-112) and rotation with kicks (gameField.cpp:55-103)
// pcn = column nibbles of the piece at its current rotation, kept in a register across keys
template <class G, int S, class ST> B2T_HD void rotate_key(ST& s, const Board<G, S>& b, const Tabs& tb, int turn, uint32_t& pcn_cur) {
    const int nr = (s.rot + turn) & 3;
    const uint32_t pcn = pcn_of(tb, s.kind, nr);
    int x = s.x, y = s.y;
    bool ok = fits<G, S>(b, x, y, pcn);
    if (!ok) {
        // kickTest walks these offsets cumulatively: (dy,dx) = (1,0)(0,-1)(0,1)(1,-1)(1,1)(0,-2)(0,2);
        // when all fail the reference undoes the net (+2,0) shift and the rotation (gameField.cpp:62-64)
        for (int k = 0; k < 7 && !ok; k++) {
            const int dy = (0x19 >> k) & 1;                          // dy = 1 for k = 0,3,4
            const int dx = (int)((0x1032CAu >> (3 * k)) & 7u) - 2;   // dx+2 = 2,1,3,1,3,0,4 packed 3 bits each
            if (fits<G, S>(b, s.x + dx, s.y + dy, pcn)) { ok = true; x = s.x + dx; y = s.y + dy; }
        }
    }
    if (ok) { s.rot = nr; s.x = x; s.y = y; pcn_cur = pcn; }
}
// `run` consecutive copies of one key (PythonHandle.cpp:73-112).  Repeats are folded where the outcome allows it: a lateral
// step that is blocked stays blocked, n single downs are one descent of min(n, free rows) with the same DropDelay
// bookkeeping (reset after any successful step, set after the first failed one; DropDelay.cpp:23-41 is idempotent).
template <class G, int S, class ST> B2T_HD void apply_key_run(ST& s, const Board<G, S>& b, const Tabs& tb, int key, int run, uint32_t& pcn) {
    if (key >= 1 && key <= 4) {
        const int dx = (key & 1) ? key - 2 : key - 3;            // 1,2 -> -1 ; 3,4 -> +1
        const int steps = (key == 2 || key == 4) ? G::W : run;    // slam: until blocked
        for (int k = 0; k < steps; k++) {
            if (!fits<G, S>(b, s.x + dx, s.y, pcn)) break;
            s.x += dx;
        }
    } else if (key == 5 || key == 6) {                            // GamePlay::mDown (gamePlay.cpp:61-69), once each / until blocked
        const int dist = rest_from<G>(collide_mask<G, S>(b, s.x, pcc_of(tb, s.kind, s.rot)), s.y) - s.y;
        const int want = (key == 6) ? dist + 1 : run;             // key 6 always ends with one failed step
        const int d = want < dist ? want : dist;
        if (d > 0) { s.y += d; dd_reset(s, s.time_ms); }
        if (d < want) dd_set(s, s.time_ms);
    } else if (key >= 8 && key <= 10) {
        for (int k = 0; k < run; k++) rotate_key<G, S>(s, b, tb, key == 8 ? 1 : key == 9 ? 3 : 2, pcn);
    } else if (key == 7) {
        hd_make<G, S>(s, b, tb);                                   // locking twice in a row changes nothing
    }
}
// The acting player's whole key string (4-bit tokens, word j at keys[j*KS]).  Every enumerated placement ends with the hard
// drop; it is peeled off the loop so that all lanes of a warp lock their piece together instead of once per distinct length,
// and the loop advances by runs of equal keys (the "n rows down" stretches of under-overhang placements are one step).
template <class G, int S, int KS, class ST>
B2T_HD void apply_keys(ST& s, const Board<G, S>& b, const Tabs& tb, const uint32_t* keys, int nkeys) {
    if (s.kind > 6 || nkeys <= 0) return;
    uint32_t pcn = pcn_of(tb, s.kind, s.rot);
    const int last = nkeys - 1;
    const bool final_drop = ((keys[(last >> 3) * KS] >> (4 * (last & 7))) & 15u) == 7u;
    const int nloop = final_drop ? last : nkeys;
    int i = 0;
    while (i < nloop) {
        const uint32_t w = keys[(i >> 3) * KS] >> (4 * (i & 7));   // this word from token i on (zeros shifted in above)
        const uint32_t key = w & 15u;
        const uint32_t diff = w ^ (key * 0x11111111u);              // zero nibbles where the token repeats
        int run = diff ? (ctz32(diff) >> 2) : 8;
        const int avail = 8 - (i & 7);
        if (run > avail) run = avail;
        if (run > nloop - i) run = nloop - i;
        apply_key_run<G, S>(s, b, tb, (int)key, run, pcn);
        i += run;
    }
    if (final_drop) hd_make<G, S>(s, b, tb);
}

// ---- PythonHandle.cpp:149-188 finish_actions for the 2-player game. gf = game flags. Returns round_over.
template <class G, int S>
B2T_HD bool finish_actions(PS& p0, const Board<G, S>& b0, PS& p1, const Board<G, S>& b1, const Tabs& tb, uint32_t seed16,
                           int elapsed, int& gf) {
    if (gf & GF_ROUND_OVER) return true;
    // phase A (:151-158): hd_finish in player order; a death ends the phase for everyone after
    bool stop = false;
    if (!p0.dead) {
        const int sent = hd_finish<G, S>(p0, b0, tb, seed16);
        if (sent == -1) { p0.dead = 1; stop = true; }
        else if (sent) p1.incoming += sent;   // distributeLines :124-136 with one opponent: amount/1.0f
    }
    if (!stop && !p1.dead) {
        const int sent = hd_finish<G, S>(p1, b1, tb, seed16);
        if (sent == -1) p1.dead = 1;
        else if (sent) p0.incoming += sent;
    }
    // phase B (:160-180)
    int alive = 0;
    if (!p0.dead) {
        const int sent = delay_check<G, S>(p0, b0, tb, seed16, elapsed);
        if (sent == -1) p0.dead = 1;
        else {
            if (sent) p1.incoming += sent;
            alive++;
            p0.reward = p0.pending & 255; p0.pending = 0;
            p0.inc_count = gq_total<S>(p0) & 255;
        }
    }
    if (!p1.dead) {
        const int sent = delay_check<G, S>(p1, b1, tb, seed16, elapsed);
        if (sent == -1) p1.dead = 1;
        else {
            if (sent) p0.incoming += sent;
            alive++;
            p1.reward = p1.pending & 255; p1.pending = 0;
            p1.inc_count = gq_total<S>(p1) & 255;
        }
    }
    if (alive < 2) { gf |= GF_ROUND_OVER; return true; }
    return false;
}

// gamePlay.cpp:206-216 restartRound (+ what it does NOT clear) and :218-230 seed via the stream table
template <class G, int S> B2T_HD void restart_and_seed(PS& s, const Board<G, S>& b, const Tabs& tb, uint32_t seed16) {
    board_clear<G, S>(b);
    gq_clear<S>(s);
    s.min_remaining = 1000;
    s.combo_start = 0; s.combo_time = 0; s.combo_count = 0; s.line_count = 0;
    s.pending = 0;
    s.drop_delay = 1000; s.drop_delay_time = 0; s.inc_dd_time = 0; s.lockdown_time = 0; s.lockdown = 0;
    s.time_ms = 0; s.incoming = 0; s.dead = 0;
    s.flags = 0;
    s.hole_ctr = 0;
    s.kind = tb.piece[(size_t)seed16 * tb.piece_len + 0];
    s.next = tb.piece[(size_t)seed16 * tb.piece_len + 1];
    s.piece_ctr = 2;
    s.rot = spawn_rot_of(tb, s.kind);
    s.x = G::SPAWN_X; s.y = 0;
}
// PythonHandle.cpp:49-66 reset
template <class G, int S>
B2T_HD void reset_game(PS& p0, const Board<G, S>& b0, PS& p1, const Board<G, S>& b1, const Tabs& tb, uint32_t seed16, int& gf) {
    int winner = -1, alive = 0;
    if (!p0.dead) { alive++; winner = 0; }
    if (!p1.dead) { alive++; winner = 1; }
    if (alive > 1) winner = -1;
    restart_and_seed<G, S>(p0, b0, tb, seed16);
    restart_and_seed<G, S>(p1, b1, tb, seed16);
    gf = ((winner + 1) << GF_WINNER_SHIFT);
}

// ------------------------------------------------------------------------------------------------ any number of players
// The two-player game above is what every BASELINE configuration and every preset of the reference runs, and what the fast
// kernels are built for.  PythonHandle itself takes any player count (PythonHandle.cpp:5-25): with one player a round ends when
// that player dies (check_for_winner stays false), with three or more the lines a player sends are split evenly — as a float —
// among ALL the others, dead or alive (distributeLines :124-136), and the fraction a player has not yet received as a whole line
// stays in GamePlay::incoming_lines (gamePlay.cpp:97-102).  These routines restate that for 1..MAX_PLAYERS players; the kernels
// that use them (gamen_kernel) keep the scalars of all players in local memory and are the compatibility path for n_players != 2.
constexpr int MAX_PLAYERS = 4;
B2T_HD float f_sub(float a, float b) {
#if defined(__CUDA_ARCH__)
    return __fsub_rn(a, b);
#else
    volatile float r = a - b; return r;
#endif
}
// gamePlay.cpp:90-114 delayCheck with the float incoming_lines of the reference: -1 = died, else lines sent
template <class G, int S> B2T_HD int delay_check_f(PS& s, float& incoming, const Board<G, S>& b, const Tabs& tb, uint32_t seed16, int elapsed) {
    s.time_ms += elapsed;
    if (dd_check(s, s.time_ms)) gp_move_down<G, S>(s, b, tb);
    if (s.lockdown && s.time_ms > s.lockdown_time && !gp_move_down<G, S>(s, b, tb)) {  // DropDelay.cpp:43-48
        hd_make<G, S>(s, b, tb);
        return hd_finish<G, S>(s, b, tb, seed16);
    }
    int add = 0;
    while (incoming >= 1.f) { add++; incoming = f_sub(incoming, 1.f); }                  // :97-101
    if (add) gq_add<S>(s, add, s.time_ms);
    int sent = cb_check(s, s.time_ms, tb);
    if (sent) sent = gq_block<S>(s, sent, s.time_ms, false);
    if (gq_check<S>(s, s.time_ms))
        if (push_garbage<G, S>(s, b, tb, seed16)) return -1;
    return sent;
}
// PythonHandle.cpp:124-136 distributeLines
B2T_HD void distribute_lines(float* incoming, int np, int sender, int amount) {
    const float others = (float)(np - 1);
    if (others < 1.f) return;
    const float each = f_div((float)amount, others);
    for (int i = 0; i < np; i++)
        if (i != sender) incoming[i] = f_add(incoming[i], each);
}
// PythonHandle.cpp:149-188 finish_actions for np players; returns round_over
template <class G, int S>
B2T_HD bool finish_actions_n(PS* p, const Board<G, S>* b, float* incoming, int np, const Tabs& tb, uint32_t seed16, int elapsed, int& gf) {
    if (gf & GF_ROUND_OVER) return true;
    for (int i = 0; i < np; i++) {                     // phase A (:151-158): a death ends it for everyone after
        if (p[i].dead) continue;
        const int sent = hd_finish<G, S>(p[i], b[i], tb, seed16);
        if (sent == -1) { p[i].dead = 1; break; }
        if (sent) distribute_lines(incoming, np, i, sent);
    }
    int alive = 0;
    for (int i = 0; i < np; i++) {                     // phase B (:160-180)
        if (p[i].dead) continue;
        const int sent = delay_check_f<G, S>(p[i], incoming[i], b[i], tb, seed16, elapsed);
        if (sent == -1) { p[i].dead = 1; continue; }
        if (sent) distribute_lines(incoming, np, i, sent);
        alive++;
        p[i].reward = p[i].pending & 255; p[i].pending = 0;
        p[i].inc_count = gq_total<S>(p[i]) & 255;
    }
    if ((np > 1 && alive < 2) || !alive) { gf |= GF_ROUND_OVER; return true; }   // check_for_winner = n_players > 1 (:13)
    return false;
}
// PythonHandle.cpp:49-66 reset for np players (last_winner: the last alive player in index order, -1 if several are alive or none;
// `if (!player_count) last_winner = 0` makes it 0 for a single player)
template <class G, int S>
B2T_HD void reset_game_n(PS* p, const Board<G, S>* b, float* incoming, int np, const Tabs& tb, uint32_t seed16, int& gf) {
    int winner = -1, alive = 0;
    for (int i = 0; i < np; i++) {
        if (!p[i].dead) { alive++; winner = i; }
        restart_and_seed<G, S>(p[i], b[i], tb, seed16);
        incoming[i] = 0.f;
    }
    if (np == 1) winner = 0;
    if (alive > 1) winner = -1;
    gf = (((winner + 1) & 3) << GF_WINNER_SHIFT);   // (two bits: not observable through the n-player entry points)
}
// tetris_environment.py:135-149 reward_fcn reads backend.states[1 - player]: a Python index, so for players beyond the second it
// counts from the end of the list (player 2 of 3 reads itself)
B2T_HD int opponent_index(int player, int np) {
    const int i = 1 - player;
    return i >= 0 ? i : np + i;
}

// Engine-chosen seed of a round (b2t_step_args.auto_reset): what time(NULL) is to the reference (PythonHandle.cpp:68-71) — any
// value will do, only its low 16 bits matter — as a counter hash of (handle seed, game id, rounds started), so a run is
// reproducible and independent of how games are batched or sharded.
B2T_HD uint32_t auto_seed(uint32_t base, uint32_t game, uint32_t episode) {
    uint32_t z = base + game * 0x9E3779B1u + episode * 0x85EBCA77u;
    z ^= z >> 15; z *= 0x2C1B3C6Du; z ^= z >> 12; z *= 0x297A2D39u; z ^= z >> 15;
    return z & 0xFFFFu;
}

// tetris_environment.py:135-149 reward_fcn, extra_rewards=False
B2T_HD int reward_base(bool done, int me_dead, int you_dead) {
    if (!done) return 0;
    if (me_dead && you_dead) return -1;
    return (you_dead ? 1 : 0) - (me_dead ? 1 : 0);
}

// state_processors.py:23-54 vector part: 19 bytes
B2T_HD void observe_vec(const PS& s, uint8_t* v) {
    v[0] = (uint8_t)s.x; v[1] = (uint8_t)s.y; v[2] = (uint8_t)s.inc_count;
    const uint32_t ct = (uint32_t)(s.remaining + 50) & 0xFFFFu;
    v[3] = (uint8_t)((ct < 25000u ? ct : 25000u) / 100u);
    v[4] = (uint8_t)s.combo_count;
#pragma unroll
    for (int k = 0; k < 7; k++) { v[5 + k] = (s.next == k); v[12 + k] = (s.kind == k); }
}

}  // namespace b2t
