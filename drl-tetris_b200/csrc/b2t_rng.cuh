// b200tetris — counter-indexed reproduction of the reference's piece and garbage-hole streams.
//
// The reference seeds two std::mt19937 per player from a value truncated to a short
// (randomizer.cpp:34-36,47-49; PythonHandle.cpp:68-71), so only 65 536 distinct streams exist, and what a
// player draws is a pure function of (seed, number of draws so far): the "bag-ish" float weights of
// randomizer.cpp:10-32 evolve deterministically.  Instead of carrying 2 x 2.5 KB of Mersenne state per player
// we tabulate, once per handle, piece[seed][k] and hole[seed][k]; a game then keeps a 16-bit seed and two
// counters.  One thread fills one seed row, holding the MT19937 state in local memory.
#pragma once
#include "b2t_core.cuh"

namespace b2t {

struct MT19937 {
    uint32_t s[624];
    int at;
};
B2T_HD void mt_seed(MT19937& m, uint32_t seed) {
    m.s[0] = seed;
    for (int i = 1; i < 624; i++) m.s[i] = 1812433253u * (m.s[i - 1] ^ (m.s[i - 1] >> 30)) + (uint32_t)i;
    m.at = 624;
}
B2T_HD uint32_t mt_mix(uint32_t a, uint32_t b, uint32_t far_) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return far_ ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
B2T_HD uint32_t mt_next(MT19937& m) {
    if (m.at >= 624) {
        for (int i = 0; i < 227; i++) m.s[i] = mt_mix(m.s[i], m.s[i + 1], m.s[i + 397]);
        for (int i = 227; i < 623; i++) m.s[i] = mt_mix(m.s[i], m.s[i + 1], m.s[i - 227]);
        m.s[623] = mt_mix(m.s[623], m.s[0], m.s[396]);
        m.at = 0;
    }
    uint32_t y = m.s[m.at++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
// randomizer.h:8-34: float( double(g) * 2^-32 ) — may round up to exactly 1.0f
B2T_HD float mt_unit(MT19937& m) {
    const double d = d_mul((double)mt_next(m), 1.0 / 4294967296.0);
#if defined(__CUDA_ARCH__)
    return __double2float_rn(d);
#else
    return (float)d;
#endif
}
B2T_HD float narrow_div6(float a) {  // randomizer.cpp:25 `adjust/=6.0` is a double division narrowed back
#if defined(__CUDA_ARCH__)
    return __double2float_rn(__ddiv_rn((double)a, 6.0));
#else
    volatile double d = (double)a / 6.0; return (float)d;
#endif
}
// randomizer.cpp:10-32 getPiece
B2T_HD int draw_piece(MT19937& m, float* wgt) {
    int pick = 0;
    float r = f_mul(mt_unit(m), 1000.f);
    for (int i = 0; i < 7; i++) {
        r = f_add(r, -wgt[i]);
        if (r < 0) { pick = i; break; }
    }
    float adjust = f_mul(f_div(wgt[pick], 4.f), 3.f);
    wgt[pick] = f_add(wgt[pick], -adjust);
    adjust = narrow_div6(adjust);
    for (int i = 0; i < 7; i++)
        if (i != pick) wgt[i] = f_add(wgt[i], adjust);
    return pick;
}
// Fill row `seed16` of both tables.  piece_map = (pieces*7)[:7] (tetris_environment.py:191-193);
// only_zs as PythonHandle.h:116-121.  mt is caller-provided scratch (local memory on the device).
B2T_HD void fill_stream_row(MT19937& mt, uint32_t seed16, const uint8_t* piece_map, bool only_zs, int width,
                            uint8_t* piece_row, uint32_t piece_len, uint8_t* hole_row, uint32_t hole_len) {
    const uint32_t s32 = (uint32_t)(int32_t)(int16_t)(uint16_t)seed16;
    float wgt[7];
    // GamePlay::seed, gamePlay.cpp:218-230: redraw (with fresh weights) while the first "next" is S or Z
    mt_seed(mt, s32);
    for (int i = 0; i < 7; i++) wgt[i] = 142.f;   // randomizer.cpp:58: 1000/7 in integers
    int d = piece_map[draw_piece(mt, wgt)];
    if (!only_zs)
        while (d == 2 || d == 3) {
            for (int i = 0; i < 7; i++) wgt[i] = 142.f;
            d = piece_map[draw_piece(mt, wgt)];
        }
    piece_row[0] = (uint8_t)d;
    for (uint32_t k = 1; k < piece_len; k++) piece_row[k] = piece_map[draw_piece(mt, wgt)];
    // randomizer.cpp:38-45 getHole: short(float(u*W))
    mt_seed(mt, s32);
    for (uint32_t k = 0; k < hole_len; k++) hole_row[k] = (uint8_t)(int16_t)(int)f_mul(mt_unit(mt), (float)width);
}

}  // namespace b2t
