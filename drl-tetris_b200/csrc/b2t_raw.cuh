// b200tetris — host-side codec between the packed HBM words of ONE game and b2t_game_raw (include/b200tetris.h): the game in
// the reference's own terms (State views of PythonHandle.h:54-82 + the hidden members its pickle support walks, :207-308).
// Used by b2t_get_game / b2t_set_game (per-env facade, pickling, white-box tests); never on the step path.
#pragma once
#include <cstring>

#include "../../include/b200tetris.h"
#include "b2t_game.cuh"

namespace b2t {

// colour code of a piece's grid cells (pieces.cpp:5-53 as dumped in SURVEY A.2; inverse of state_processors.py:24 col_code)
inline int piece_colour(int kind) {
    static const uint8_t colour[8] = {4, 3, 5, 7, 2, 1, 6, 0};
    return colour[kind & 7];
}

template <class G> struct HostGame {   // a Game<G,1> bound to plain host arrays
    typedef typename G::col_t col_t;
    Game<G, 1> g;
    col_t cols[2][G::NCOL];
    uint32_t stream[2][G::SW];
    uint32_t gqw[2][GQ_WORDS];
    HostGame() {
        for (int p = 0; p < 2; p++) {
            g.b[p].cols = cols[p]; g.b[p].stream = stream[p]; g.p[p].gq = gqw[p];
            board_init_walls<G, 1>(g.b[p]);
        }
    }
    void load(const uint32_t* words) {
        for (int i = 0; i < 4 * G::NV; i++) game_put_word(g, i, words[i]);
        game_unpack(g);
    }
    void store(uint32_t* words) {
        game_pack(g);
        for (int i = 0; i < 4 * G::NV; i++) words[i] = game_get_word(g, i);
    }
};

// planes: the game's 2 x TILE_PLANES x SW colour-plane words (b2t_config.flags & B2T_FLAG_TILES) or NULL
template <class G> void decode_game(const uint32_t* words, b2t_game_raw* out, uint8_t* field, const uint32_t* planes) {
    HostGame<G> hg;
    hg.load(words);
    PieceTables pt;
    make_piece_tables(pt);
    memset(out, 0, sizeof *out);
    const int gf = hg.g.gf();
    out->seed16 = (int32_t)hg.g.seed16();
    out->round_over = (gf & GF_ROUND_OVER) ? 1 : 0;
    out->done = (gf & GF_ENV_DONE) ? 1 : 0;
    out->last_winner = ((gf >> GF_WINNER_SHIFT) & 3) - 1;
    out->rounds_started = (int32_t)hg.g.p[0].episode;
    for (int p = 0; p < 2; p++) {
        const PS& s = hg.g.p[p];
        b2t_player_raw& r = out->player[p];
        if (s.kind <= 6) {
            const uint32_t pcn = pt.pcn[s.kind * 4 + s.rot];
            for (int c = 0; c < 4; c++)
                for (int y = 0; y < 4; y++)
                    if ((pcn >> (4 * c + y)) & 1u) r.piece[4 * y + c] = (uint8_t)piece_colour(s.kind);
        }
        r.x = s.x; r.y = s.y; r.inc_lines = s.inc_count; r.combo_time = s.remaining; r.combo_count = s.combo_count;
        r.nextpiece = s.next; r.reward = s.reward; r.dead = s.dead;
        r.piece_kind = s.kind; r.rotation = s.rot; r.time_ms = s.time_ms;
        r.drop_delay = s.drop_delay; r.drop_delay_time = s.drop_delay_time; r.increase_drop_delay_time = s.inc_dd_time;
        r.lockdown_time = s.lockdown_time; r.lockdown = s.lockdown;
        r.combo_start = s.combo_start; r.combo_time_total = s.combo_time; r.line_count = s.line_count;
        r.min_remaining = s.min_remaining; r.incoming_lines = s.incoming; r.lines_pending = s.pending;
        r.garbage_n = s.gq_n;
        for (int i = 0; i < GQ_CAP; i++) { r.garbage_count[i] = gq_count_at<1>(s, i); r.garbage_delay[i] = gq_delay_at<1>(s, i); }
        r.piece_draws = (int32_t)s.piece_ctr; r.hole_draws = (int32_t)s.hole_ctr;
        out->overflow |= (int32_t)s.flags;
        if (field)
            for (int c = 0; c < G::H * G::W; c++) {
                uint32_t v = (hg.g.b[p].sw(c >> 5) >> (c & 31)) & 1u;
                if (v && planes)      // raw cell value 1..8 (gameField.cpp:105-110, gamePlay.cpp:202)
                    for (int k = 0; k < TILE_PLANES; k++) v += ((planes[(p * TILE_PLANES + k) * G::SW + (c >> 5)] >> (c & 31)) & 1u) << k;
                field[p * G::H * G::W + c] = (uint8_t)v;
            }
    }
}

// words: the game's current packed state (boards are kept when field == NULL), rewritten in place
template <class G> void encode_game(uint32_t* words, const b2t_game_raw* in, const uint8_t* field, uint32_t* planes) {
    typedef typename G::col_t col_t;
    HostGame<G> hg;
    hg.load(words);
    int gf = 0;
    if (in->round_over) gf |= GF_ROUND_OVER;
    if (in->done) gf |= GF_ENV_DONE;
    gf |= ((in->last_winner + 1) & 3) << GF_WINNER_SHIFT;
    for (int p = 0; p < 2; p++) {
        PS& s = hg.g.p[p];
        const b2t_player_raw& r = in->player[p];
        s.kind = (uint32_t)r.piece_kind & 7u; s.rot = (uint32_t)r.rotation & 3u; s.next = (uint32_t)r.nextpiece & 7u;
        s.x = r.x; s.y = r.y; s.dead = r.dead ? 1 : 0; s.lockdown = r.lockdown ? 1 : 0;
        s.reward = (uint32_t)r.reward & 255u; s.inc_count = (uint32_t)r.inc_lines & 255u; s.combo_count = (uint32_t)r.combo_count & 255u;
        s.line_count = (uint32_t)r.line_count & 255u; s.remaining = (uint32_t)r.combo_time & 0xFFFFu; s.pending = (uint32_t)r.lines_pending & 0xFFFFu;
        s.time_ms = r.time_ms; s.combo_start = r.combo_start; s.combo_time = r.combo_time_total;
        s.drop_delay = (uint32_t)r.drop_delay & 0xFFFFu; s.incoming = (uint32_t)r.incoming_lines & 0xFFFFu;
        s.drop_delay_time = r.drop_delay_time; s.inc_dd_time = r.increase_drop_delay_time; s.lockdown_time = r.lockdown_time;
        s.min_remaining = r.min_remaining; s.piece_ctr = (uint32_t)r.piece_draws; s.hole_ctr = (uint32_t)r.hole_draws & 0xFFFFu;
        s.flags = (uint32_t)in->overflow & 3u;
        gq_clear<1>(s);
        int n = r.garbage_n < 0 ? 0 : (r.garbage_n > GQ_CAP ? GQ_CAP : r.garbage_n);
        for (int i = 0; i < n; i++) { gq_set_count<1>(s, i, r.garbage_count[i]); gq_set_delay<1>(s, i, r.garbage_delay[i]); }
        s.gq_n = (uint32_t)n;
        if (field) {
            board_clear<G, 1>(hg.g.b[p]);
            if (planes)
                for (int w = 0; w < TILE_PLANES * G::SW; w++) planes[p * TILE_PLANES * G::SW + w] = 0u;
            for (int y = 0; y < G::H; y++)
                for (int x = 0; x < G::W; x++)
                    if (const uint32_t v = field[(p * G::H + y) * G::W + x]) {
                        hg.g.b[p].col(x) |= (col_t)(G::one() << (y + G::YB));
                        const int bit = y * G::W + x;
                        hg.g.b[p].sw(bit >> 5) |= 1u << (bit & 31);
                        if (planes)
                            for (int k = 0; k < TILE_PLANES; k++)
                                if ((((v > 8u ? 8u : v) - 1u) >> k) & 1u) planes[(p * TILE_PLANES + k) * G::SW + (bit >> 5)] |= 1u << (bit & 31);
                    }
        }
    }
    hg.g.p[0].hi12 = (uint32_t)in->seed16 & 0xFFFFu;
    hg.g.p[0].episode = (uint32_t)in->rounds_started & 0xFFFFu;
    hg.g.p[1].episode = 0;
    hg.g.set_gf(gf);
    hg.store(words);
}

}  // namespace b2t
