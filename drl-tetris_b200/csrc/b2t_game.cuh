// b200tetris — one game = two players: HBM word layout, register/shared-memory working form, and the
// perform_action body shared by the step and expand kernels (and by tests/hostsim on the CPU).
//
// HBM layout: tiles of 32 games, struct-of-vectors inside a tile:  uint4 state[n_tiles][NV][32].  Thread g's loads
// coalesce (a warp reads 512 contiguous bytes per vector) AND a warp's NV vectors are one contiguous NV*512-byte run, so a
// step streams whole DRAM pages and touches one TLB entry per warp instead of NV.  The 4*NV words of game g are
// [player0 | player1 | pad],  each player = W*CW column words, SW stream words, 24 scalars.
#pragma once
#include "b2t_core.cuh"
#include "b2t_enum.cuh"

namespace b2t {

// index (in uint4 units) of vector v of game g
template <class G> B2T_HD size_t state_index(int g, int v) { return ((size_t)(g >> 5) * G::NV + (size_t)v) * 32 + (size_t)(g & 31); }
template <class G> B2T_HD size_t state_vectors(int n_games) { return (size_t)((n_games + 31) >> 5) * G::NV * 32; }

template <class G, int S> struct Game {
    PS p[2];
    Board<G, S> b[2];
    uint32_t sc[2][SCALAR_WORDS];   // packed scalar words in flight between HBM and PS
    B2T_HD int gf() const { return (int)p[1].hi12; }   // game flags
    B2T_HD void set_gf(int v) { p[1].hi12 = (uint32_t)v & 0xFFFFu; }
    B2T_HD uint32_t seed16() const { return (uint32_t)p[0].hi12 & 0xFFFFu; }
};

// route word i (compile-time after unrolling) of a game into shared memory / the scalar staging array
template <class G, int S> B2T_HD void game_put_word(Game<G, S>& g, int i, uint32_t w) {
    const int pl = i / G::PW, k = i % G::PW;
    if (pl >= 2) return;
    if (k >= G::W * G::CW + G::SW + 14 && k < G::W * G::CW + G::SW + 14 + GQ_WORDS) {   // garbage queue: straight to its words
        g.p[pl].gq[(k - (G::W * G::CW + G::SW + 14)) * S] = w;
        return;
    }
    if (k < G::W * G::CW) {
        if (G::CW == 1) g.b[pl].col(k) = (typename G::col_t)w;
        else reinterpret_cast<uint32_t*>(&g.b[pl].col(k >> 1))[k & 1] = w;
    } else if (k < G::W * G::CW + G::SW) g.b[pl].sw(k - G::W * G::CW) = w;
    else g.sc[pl][k - G::W * G::CW - G::SW] = w;
}
template <class G, int S> B2T_HD uint32_t game_get_word(const Game<G, S>& g, int i) {
    const int pl = i / G::PW, k = i % G::PW;
    if (pl >= 2) return 0u;
    if (k >= G::W * G::CW + G::SW + 14 && k < G::W * G::CW + G::SW + 14 + GQ_WORDS)
        return g.p[pl].gq[(k - (G::W * G::CW + G::SW + 14)) * S];
    if (k < G::W * G::CW) {
        if (G::CW == 1) return (uint32_t)g.b[pl].col(k);
        return reinterpret_cast<const uint32_t*>(&g.b[pl].col(k >> 1))[k & 1];
    }
    if (k < G::W * G::CW + G::SW) return g.b[pl].sw(k - G::W * G::CW);
    return g.sc[pl][k - G::W * G::CW - G::SW];
}
template <class G, int S> B2T_HD void game_unpack(Game<G, S>& g) {   // keeps the queue pointers set by the caller
    uint32_t* q0 = g.p[0].gq; uint32_t* q1 = g.p[1].gq;
    ps_unpack(g.p[0], g.sc[0]); ps_unpack(g.p[1], g.sc[1]);
    g.p[0].gq = q0; g.p[1].gq = q1;
}
template <class G, int S> B2T_HD void game_pack(Game<G, S>& g) { ps_pack(g.p[0], g.sc[0]); ps_pack(g.p[1], g.sc[1]); }

// A never-reset game (b2t_create): empty boards, no piece, everybody alive
template <class G, int S> B2T_HD void game_init(Game<G, S>& g) {
    for (int pl = 0; pl < 2; pl++) {
        board_clear<G, S>(g.b[pl]);
        for (int k = 0; k < SCALAR_WORDS; k++) g.sc[pl][k] = 0;
        for (int j = 0; j < GQ_WORDS; j++) g.p[pl].gq[j * S] = 0u;
    }
    game_unpack(g);
    for (int pl = 0; pl < 2; pl++) { g.p[pl].kind = 7; g.p[pl].min_remaining = 1000; g.p[pl].drop_delay = 1000; }
}

// tetris_environment.perform_action (tetris_environment.py:102-116) on one game.
//   keys: packed 4-bit tokens (word j at keys[j*KS]), nkeys of them.  Returns env.done; *reward = reward_fcn base term.
template <class G, int S, int KS>
B2T_HD bool game_perform(Game<G, S>& g, const Tabs& tb, int player, const uint32_t* keys, int nkeys, bool finalize, int elapsed,
                         int* reward) {
    int gf = g.gf();
    // make_actions, PythonHandle.cpp:138-147 (the other player gets the null action)
    if (!(gf & GF_ROUND_OVER)) {
        // Only the fields a key can touch are moved into a neutral copy (selects, no dynamic register indexing);
        // the board is selected by pointer.
        const bool one = player != 0;
        Cur me;
        me.kind = one ? g.p[1].kind : g.p[0].kind;
        me.rot = one ? g.p[1].rot : g.p[0].rot;
        me.x = one ? g.p[1].x : g.p[0].x;
        me.y = one ? g.p[1].y : g.p[0].y;
        me.time_ms = one ? g.p[1].time_ms : g.p[0].time_ms;
        me.drop_delay_time = one ? g.p[1].drop_delay_time : g.p[0].drop_delay_time;
        me.lockdown = one ? g.p[1].lockdown : g.p[0].lockdown;
        me.lockdown_time = one ? g.p[1].lockdown_time : g.p[0].lockdown_time;
        const bool dead = (one ? g.p[1].dead : g.p[0].dead) != 0;
        Board<G, S> bb;
        bb.cols = one ? g.b[1].cols : g.b[0].cols;
        bb.stream = one ? g.b[1].stream : g.b[0].stream;
        bb.tiles = one ? g.b[1].tiles : g.b[0].tiles;
        if (!dead) {
            apply_keys<G, S, KS>(me, bb, tb, keys, nkeys);
#pragma unroll
            for (int q = 0; q < 2; q++)
                if ((q == 1) == one) {
                    g.p[q].rot = me.rot; g.p[q].x = me.x; g.p[q].y = me.y;
                    g.p[q].drop_delay_time = me.drop_delay_time; g.p[q].lockdown = me.lockdown;
                    g.p[q].lockdown_time = me.lockdown_time;
                }
        }
    }
    if (finalize) {
        const bool over = finish_actions<G, S>(g.p[0], g.b[0], g.p[1], g.b[1], tb, g.seed16(), elapsed, gf);
        gf = over ? (gf | GF_ENV_DONE) : (gf & ~GF_ENV_DONE);
    }
    g.set_gf(gf);
    const bool done = (gf & GF_ENV_DONE) != 0;
    *reward = player ? reward_base(done, g.p[1].dead, g.p[0].dead) : reward_base(done, g.p[0].dead, g.p[1].dead);
    return done;
}

template <class G, int S> B2T_HD void game_reset(Game<G, S>& g, const Tabs& tb, uint32_t seed) {
    int gf = g.gf();
    const uint32_t s16 = seed & 0xFFFFu;
    reset_game<G, S>(g.p[0], g.b[0], g.p[1], g.b[1], tb, s16, gf);
    g.p[0].episode = (g.p[0].episode + 1u) & 0xFFFFu;
    g.p[0].hi12 = s16;
    g.p[1].hi12 = (uint32_t)gf & 0xFFFFu;   // round_over and env_done cleared (tetris_environment.py:66-68)
}

// Convert u8 keys (any value; >10 are no-ops like PythonHandle.cpp:109-110) to packed tokens
template <int KW> B2T_HD void pack_keys(const uint8_t* keys, int n, uint32_t* out) {
#pragma unroll
    for (int j = 0; j < KW; j++) {   // static word / nibble indices only (out[] stays in registers)
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const int i = 8 * j + b;
            if (i < n) {
                const uint32_t k = keys[i];
                w |= (k <= 10u ? k : 0u) << (4 * b);
            }
        }
        out[j] = w;
    }
}
template <int KW> B2T_HD int packed_len(const uint32_t* w) {  // sequences hold no zero token except the null action
    int n = 1;
    for (int i = 0; i < KW; i++)
        if (w[i]) n = i * 8 + (msb32(w[i]) >> 2) + 1;
    return n;
}

}  // namespace b2t
