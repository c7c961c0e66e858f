// TEST INFRASTRUCTURE ONLY.  Runs the *same* game-core headers the CUDA kernels are built from
// (drl-tetris_b200/csrc/b2t_{core,enum,game,rng}.cuh) on the CPU, one game at a time, through the identical HBM
// word layout, so rule bugs are found in the GPU-less build container by diffing against the oracle
// (tests/test_hostsim_vs_oracle.py).  Nothing in the product library links or calls this.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "b2t_game.cuh"
#include "b2t_raw.cuh"
#include "b2t_geoms.h"
#include "b2t_rng.cuh"

using namespace b2t;

#ifdef B2T_COUNTERS
extern "C" { long long b2t_counters[16] = {0}; }
#endif

struct HSBase {
    virtual ~HSBase() {}
    virtual void reset(const int32_t* idx, int n, const uint32_t* seeds) = 0;
    virtual void step(const int32_t* idx, int n, const uint8_t* player, const uint8_t* keys, const uint8_t* key_len, int lmax,
                      int time_ms, int finalize, int8_t* reward, uint8_t* done, uint8_t* lines) = 0;
    virtual void observe(const int32_t* idx, int n, uint8_t* field, uint8_t* vec, uint8_t* pidx) = 0;
    virtual void scal(const int32_t* idx, int n, int32_t* out) = 0;
    virtual int enumerate(const int32_t* idx, int n, const uint8_t* player, int bar_null, int kmax, int lmax, int32_t* K,
                          uint8_t* keys, uint8_t* key_len) = 0;
    virtual int expand(const int32_t* idx, int n, const uint8_t* player, int bar_null, int finalize, int time_ms, int kmax, int lmax,
                       int32_t* K, uint8_t* keys, uint8_t* key_len, uint8_t* field, uint8_t* vec, uint8_t* pidx, int8_t* reward,
                       uint8_t* done, uint8_t* lines, uint8_t* dead) = 0;
    virtual int flags(int game) = 0;
    virtual void get_game(int game, b2t_game_raw* out, uint8_t* field) = 0;
    virtual void set_game(int game, const b2t_game_raw* in, const uint8_t* field) = 0;
    virtual void copy_game(int dst, HSBase* from, int src) = 0;
    virtual void words_of(int game, uint32_t* out) = 0;
    virtual int nwords() = 0;
    virtual void enable_tiles() = 0;
    virtual uint32_t* planes_of(int game) = 0;
    virtual int blob_bytes() { return 0; }                       // opaque per-game snapshot (n-player simulator only)
    virtual void get_blob(int, void*) {}
    virtual void set_blob(int, const void*) {}
};

template <int KW> struct ListSink {
    uint32_t* out;  // [cap][KW]
    int cap, count, bad;
    uint64_t fin[4];
    void init(uint32_t* o, int c) { out = o; cap = c; count = 0; bad = 0; fin[0] = fin[1] = fin[2] = fin[3] = 0; }
    B2T_HD void append(const Seq<KW>& s, bool b, bool is_fin) {
        if (b) bad = 1;
        if (count >= cap) { bad = 1; return; }
        for (int i = 0; i < KW; i++) out[count * KW + i] = s.w[i];
        if (is_fin) fin[count >> 6] |= 1ull << (count & 63);
        count++;
    }
    B2T_HD void top(const Seq<KW>& s, bool b) { append(s, b, false); }
    B2T_HD void finesse(const Seq<KW>& s, bool b) {
        for (int k = 0; k < count; k++)
            if ((fin[k >> 6] >> (k & 63)) & 1)
                if (seq_equal<KW>(s, out + k * KW)) return;   // action_list.py:14: first occurrence wins
        append(s, b, true);
    }
};

// piece / hole stream tables (filled lazily per seed), piece shapes, the combo power table: what DevTabs is on the device
template <int W> struct HSTabs {
    PieceTables pt;
    uint8_t map[7];
    bool only_zs;
    uint32_t piece_len, hole_len;
    std::vector<uint8_t> piece_tab, hole_tab;
    std::vector<uint8_t> row_ready;
    double powtab[256];
    Tabs tb;
    HSTabs(const uint8_t* map_, uint32_t plen, uint32_t hlen) : piece_len(plen), hole_len(hlen) {
        memcpy(map, map_, 7);
        only_zs = true;
        for (int i = 0; i < 7; i++) if (map[i] != 2 && map[i] != 3) only_zs = false;
        make_piece_tables(pt);
        piece_tab.assign((size_t)RNG_SEEDS * piece_len, 0);
        hole_tab.assign((size_t)RNG_SEEDS * hole_len, 0);
        row_ready.assign(RNG_SEEDS, 0);
        for (int c = 0; c < 256; c++) powtab[c] = pow((double)c, 1.4 + c * 0.01);
        tb.pcn = pt.pcn; tb.pcc = pt.pcc; tb.piece = piece_tab.data(); tb.hole = hole_tab.data(); tb.powtab = powtab;
        tb.piece_len = piece_len; tb.hole_len = hole_len; tb.spawn_rots = 0;
        for (int k = 0; k < 7; k++) tb.spawn_rots |= (uint32_t)pt.spawn_rot[k] << (2 * k);
    }
    void ensure_row(uint32_t s16) {
        if (row_ready[s16]) return;
        static MT19937 mt;
        fill_stream_row(mt, s16, map, only_zs, W, &piece_tab[(size_t)s16 * piece_len], piece_len, &hole_tab[(size_t)s16 * hole_len], hole_len);
        row_ready[s16] = 1;
    }
};

template <int H, int W> struct HS : HSBase, HSTabs<W> {
    typedef Geo<H, W> G;
    typedef typename G::col_t col_t;
    static constexpr int KW = G::KEYW;
    using HSTabs<W>::tb;
    using HSTabs<W>::ensure_row;
    int n;
    std::vector<uint32_t> state;  // uint4 state[NV][n]
    col_t cols[2][G::NCOL];
    uint32_t stream[2][G::SW];
    uint32_t gqw[2][GQ_WORDS];
    std::vector<uint32_t> tiles;   // optional colour planes [n][2][TILE_PLANES][SW] (B2T_FLAG_TILES of the product)
    static constexpr int TW = 2 * TILE_PLANES * G::SW;
    void enable_tiles() override { tiles.assign((size_t)n * TW, 0u); }
    uint32_t* planes_of(int gi) override { return tiles.empty() ? nullptr : tiles.data() + (size_t)gi * TW; }

    HS(int n_, const uint8_t* map_, uint32_t plen, uint32_t hlen) : HSTabs<W>(map_, plen, hlen), n(n_) {
        state.assign(state_vectors<G>(n) * 4, 0);
        Game<G, 1> g;
        bind(g);
        for (int i = 0; i < n; i++) { game_init(g); save(g, i); }
    }
    void bind(Game<G, 1>& g) {
        for (int p = 0; p < 2; p++) {
            g.b[p].cols = cols[p]; g.b[p].stream = stream[p]; g.p[p].gq = gqw[p];
            board_init_walls<G, 1>(g.b[p]);
        }
    }
    void load(Game<G, 1>& g, int gi) {
        for (int p = 0; p < 2; p++) g.b[p].tiles = tiles.empty() ? nullptr : planes_of(gi) + p * TILE_PLANES * G::SW;
        for (int v = 0; v < G::NV; v++)
            for (int k = 0; k < 4; k++) game_put_word(g, 4 * v + k, state[state_index<G>(gi, v) * 4 + k]);
        game_unpack(g);
    }
    void save(Game<G, 1>& g, int gi) {
        game_pack(g);
        for (int v = 0; v < G::NV; v++)
            for (int k = 0; k < 4; k++) state[state_index<G>(gi, v) * 4 + k] = game_get_word(g, 4 * v + k);
    }
    void reset(const int32_t* idx, int cnt, const uint32_t* seeds) override {
        Game<G, 1> g; bind(g);
        for (int i = 0; i < cnt; i++) {
            int gi = idx ? idx[i] : i;
            ensure_row(seeds[i] & 0xFFFF);
            load(g, gi);
            game_reset(g, tb, seeds[i]);
            save(g, gi);
        }
    }
    void step(const int32_t* idx, int cnt, const uint8_t* player, const uint8_t* keys, const uint8_t* key_len, int lmax, int time_ms,
              int finalize, int8_t* reward, uint8_t* done, uint8_t* lines) override {
        Game<G, 1> g; bind(g);
        for (int i = 0; i < cnt; i++) {
            int gi = idx ? idx[i] : i;
            load(g, gi);
            uint32_t kw[KW];
            pack_keys<KW>(keys + (size_t)i * lmax, key_len[i], kw);
            int r;
            bool d = game_perform<G, 1, 1>(g, tb, player[i], kw, key_len[i], finalize != 0, time_ms, &r);
            if (reward) reward[i] = (int8_t)r;
            if (done) done[i] = d;
            if (lines) { lines[2 * i] = (uint8_t)g.p[0].reward; lines[2 * i + 1] = (uint8_t)g.p[1].reward; }
            save(g, gi);
        }
    }
    void obs_of(Game<G, 1>& g, uint8_t* field, uint8_t* vec, uint8_t* pidx) {
        for (int p = 0; p < 2; p++) {
            if (field)
                for (int c = 0; c < H * W; c++) field[p * H * W + c] = (g.b[p].sw(c >> 5) >> (c & 31)) & 1;
            if (vec) observe_vec(g.p[p], vec + p * VEC_LEN);
            if (pidx) pidx[p] = (uint8_t)g.p[p].kind;
        }
        // cross-check the two board forms (columns vs stream)
        for (int p = 0; p < 2; p++)
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) {
                    int a = (int)((g.b[p].col(x) >> (y + G::YB)) & 1), b = (g.b[p].sw((y * W + x) >> 5) >> ((y * W + x) & 31)) & 1;
                    if (a != b) abort();
                }
    }
    void observe(const int32_t* idx, int cnt, uint8_t* field, uint8_t* vec, uint8_t* pidx) override {
        Game<G, 1> g; bind(g);
        for (int i = 0; i < cnt; i++) {
            load(g, idx ? idx[i] : i);
            obs_of(g, field ? field + (size_t)i * 2 * H * W : nullptr, vec ? vec + (size_t)i * 2 * VEC_LEN : nullptr, pidx ? pidx + 2 * i : nullptr);
        }
    }
    void scal(const int32_t* idx, int cnt, int32_t* out) override {
        Game<G, 1> g; bind(g);
        for (int i = 0; i < cnt; i++) {
            load(g, idx ? idx[i] : i);
            for (int p = 0; p < 2; p++) {
                int32_t* s = out + ((size_t)i * 2 + p) * 8;
                const PS& q = g.p[p];
                s[0] = q.x; s[1] = q.y; s[2] = q.inc_count; s[3] = q.remaining; s[4] = q.combo_count; s[5] = q.next; s[6] = q.reward; s[7] = q.dead;
            }
        }
    }
    // enumerate one loaded game into packed list; returns K (after null handling), list at `lst` (cap KMAX+1)
    int enum_game(Game<G, 1>& g, int player, int bar_null, uint32_t* lst, int cap, int* bad) {
        ListSink<KW> sink;
        int off = bar_null ? 0 : 1;
        for (int i = 0; i < KW; i++) lst[i] = 0;   // slot 0 = null action when needed
        sink.init(lst + off * KW, cap - off);
        const PS& me = g.p[player];
        if (me.kind <= 6) {
            uint8_t pbuf[EnumGeo<G>::PCAP];
            Enumerator<G, 1, KW, ListSink<KW>> e(g.b[player], tb, me.kind, sink, pbuf);
            e.run(me.rot);
        }
        *bad = sink.bad;
        if (sink.count == 0) { for (int i = 0; i < KW; i++) lst[i] = 0; return 1; }
        return sink.count + off;
    }
    static void unpack_seq(const uint32_t* w, uint8_t* keys, uint8_t* len, int lmax) {
        int L = packed_len<KW>(w);
        memset(keys, 0, lmax);
        for (int i = 0; i < L && i < lmax; i++) keys[i] = (w[i >> 3] >> (4 * (i & 7))) & 15;
        *len = (uint8_t)(L < lmax ? L : lmax);
    }
    int enumerate(const int32_t* idx, int cnt, const uint8_t* player, int bar_null, int kmax, int lmax, int32_t* K, uint8_t* keys,
                  uint8_t* key_len) override {
        Game<G, 1> g; bind(g);
        std::vector<uint32_t> lst((size_t)(kmax + 1) * KW);
        int nbad = 0;
        for (int i = 0; i < cnt; i++) {
            load(g, idx ? idx[i] : i);
            int bad;
            int k = enum_game(g, player[i], bar_null, lst.data(), kmax, &bad);
            K[i] = k;
            nbad += bad;
            for (int a = 0; a < kmax; a++) {
                uint8_t* dst = keys + ((size_t)i * kmax + a) * lmax;
                if (a < k) unpack_seq(&lst[(size_t)a * KW], dst, &key_len[(size_t)i * kmax + a], lmax);
                else { memset(dst, 0, lmax); key_len[(size_t)i * kmax + a] = 0; }
            }
        }
        return nbad;
    }
    int expand(const int32_t* idx, int cnt, const uint8_t* player, int bar_null, int finalize, int time_ms, int kmax, int lmax, int32_t* K,
               uint8_t* keys, uint8_t* key_len, uint8_t* field, uint8_t* vec, uint8_t* pidx, int8_t* reward, uint8_t* done,
               uint8_t* lines, uint8_t* dead) override {
        Game<G, 1> g; bind(g);
        std::vector<uint32_t> lst((size_t)(kmax + 1) * KW);
        int nbad = 0;
        for (int i = 0; i < cnt; i++) {
            int gi = idx ? idx[i] : i;
            load(g, gi);
            int bad;
            int k = enum_game(g, player[i], bar_null, lst.data(), kmax, &bad);
            K[i] = k; nbad += bad;
            for (int a = 0; a < kmax; a++) {
                size_t o = (size_t)i * kmax + a;
                uint8_t* dst = keys + o * lmax;
                if (a >= k) { memset(dst, 0, lmax); key_len[o] = 0; continue; }
                unpack_seq(&lst[(size_t)a * KW], dst, &key_len[o], lmax);
                load(g, gi);
                int r;
                bool d = game_perform<G, 1, 1>(g, tb, player[i], &lst[(size_t)a * KW], packed_len<KW>(&lst[(size_t)a * KW]), finalize != 0, time_ms, &r);
                if (reward) reward[o] = (int8_t)r;
                if (done) done[o] = d;
                if (lines) { lines[2 * o] = (uint8_t)g.p[0].reward; lines[2 * o + 1] = (uint8_t)g.p[1].reward; }
                if (dead) { dead[2 * o] = (uint8_t)g.p[0].dead; dead[2 * o + 1] = (uint8_t)g.p[1].dead; }
                obs_of(g, field ? field + o * 2 * H * W : nullptr, vec ? vec + o * 2 * VEC_LEN : nullptr, pidx ? pidx + 2 * o : nullptr);
            }
        }
        return nbad;
    }
    int nwords() override { return 4 * G::NV; }
    void words_of(int gi, uint32_t* out) override {
        for (int v = 0; v < G::NV; v++)
            for (int k = 0; k < 4; k++) out[4 * v + k] = state[state_index<G>(gi, v) * 4 + k];
    }
    void put_words(int gi, const uint32_t* in) {
        for (int v = 0; v < G::NV; v++)
            for (int k = 0; k < 4; k++) state[state_index<G>(gi, v) * 4 + k] = in[4 * v + k];
    }
    void get_game(int gi, b2t_game_raw* out, uint8_t* field) override {   // the product's host codec (b2t_raw.cuh) on this store
        uint32_t w[4 * G::NV];
        words_of(gi, w);
        decode_game<G>(w, out, field, planes_of(gi));
    }
    void set_game(int gi, const b2t_game_raw* in, const uint8_t* field) override {
        uint32_t w[4 * G::NV];
        words_of(gi, w);
        encode_game<G>(w, in, field, field ? planes_of(gi) : nullptr);
        put_words(gi, w);
        ensure_row((uint32_t)in->seed16 & 0xFFFF);
    }
    void copy_game(int dst, HSBase* from, int src) override {
        uint32_t w[4 * G::NV];
        from->words_of(src, w);
        put_words(dst, w);
        if (uint32_t* d = planes_of(dst)) {
            const uint32_t* sp = from->planes_of(src);
            for (int k = 0; k < TW; k++) d[k] = sp ? sp[k] : 0u;
        }
        ensure_row(seed_of(w));
    }
    uint32_t seed_of(const uint32_t* w) {
        b2t_game_raw r;
        decode_game<G>(w, &r, nullptr, nullptr);
        return (uint32_t)r.seed16 & 0xFFFF;
    }
    int flags(int gi) override {
        Game<G, 1> g; bind(g);
        load(g, gi);
        return g.p[0].flags | g.p[1].flags;
    }
};

// n_players != 2: the rules of b2t_core.cuh's "any number of players" section (what gamen_kernel runs on the device), games kept
// unpacked.  The per-game raw codec, colour planes and word views describe two players and are not offered here.
template <int H, int W> struct HSN : HSBase, HSTabs<W> {
    typedef Geo<H, W> G;
    typedef typename G::col_t col_t;
    static constexpr int KW = G::KEYW;
    using HSTabs<W>::tb;
    using HSTabs<W>::ensure_row;
    struct GameN {
        PS p[MAX_PLAYERS];
        col_t cols[MAX_PLAYERS][G::NCOL];
        uint32_t stream[MAX_PLAYERS][G::SW];
        uint32_t gq[MAX_PLAYERS][GQ_WORDS];
        float inc[MAX_PLAYERS];
        int gf;
        uint32_t seed16;
    };
    int n, np;
    std::vector<GameN> games;
    HSN(int n_, int np_, const uint8_t* map_, uint32_t plen, uint32_t hlen) : HSTabs<W>(map_, plen, hlen), n(n_), np(np_), games(n_) {
        for (auto& g : games) {
            memset(&g, 0, sizeof g);
            Board<G, 1> bd[MAX_PLAYERS];
            bind(g, bd);
            for (int pl = 0; pl < np; pl++) {
                board_init_walls<G, 1>(bd[pl]);
                board_clear<G, 1>(bd[pl]);
                g.p[pl].kind = 7; g.p[pl].min_remaining = 1000; g.p[pl].drop_delay = 1000;
            }
        }
    }
    void bind(GameN& g, Board<G, 1>* bd) {
        for (int pl = 0; pl < np; pl++) { bd[pl].cols = g.cols[pl]; bd[pl].stream = g.stream[pl]; bd[pl].tiles = nullptr; g.p[pl].gq = g.gq[pl]; }
    }
    void reset(const int32_t* idx, int cnt, const uint32_t* seeds) override {
        for (int i = 0; i < cnt; i++) {
            GameN& g = games[idx ? idx[i] : i];
            Board<G, 1> bd[MAX_PLAYERS];
            bind(g, bd);
            ensure_row(seeds[i] & 0xFFFF);
            g.seed16 = seeds[i] & 0xFFFF;
            reset_game_n<G, 1>(g.p, bd, g.inc, np, tb, g.seed16, g.gf);
        }
    }
    bool perform(GameN& g, int player, const uint32_t* kw, int nkeys, bool finalize, int time_ms, int* reward) {
        Board<G, 1> bd[MAX_PLAYERS];
        bind(g, bd);
        if (player >= np) player = np - 1;
        if (!(g.gf & GF_ROUND_OVER) && !g.p[player].dead) apply_keys<G, 1, 1>(g.p[player], bd[player], tb, kw, nkeys);
        if (finalize) {
            const bool over = finish_actions_n<G, 1>(g.p, bd, g.inc, np, tb, g.seed16, time_ms, g.gf);
            g.gf = over ? (g.gf | GF_ENV_DONE) : (g.gf & ~GF_ENV_DONE);
        }
        const bool done = (g.gf & GF_ENV_DONE) != 0;
        const bool you_dead = np > 1 && g.p[opponent_index(player, np)].dead;
        *reward = reward_base(done, g.p[player].dead, you_dead);
        return done;
    }
    void step(const int32_t* idx, int cnt, const uint8_t* player, const uint8_t* keys, const uint8_t* key_len, int lmax, int time_ms,
              int finalize, int8_t* reward, uint8_t* done, uint8_t* lines) override {
        for (int i = 0; i < cnt; i++) {
            GameN& g = games[idx ? idx[i] : i];
            uint32_t kw[KW];
            pack_keys<KW>(keys + (size_t)i * lmax, key_len[i], kw);
            int r;
            const bool d = perform(g, player[i], kw, key_len[i], finalize != 0, time_ms, &r);
            if (reward) reward[i] = (int8_t)r;
            if (done) done[i] = d;
            if (lines) for (int pl = 0; pl < np; pl++) lines[(size_t)i * np + pl] = (uint8_t)g.p[pl].reward;
        }
    }
    void obs_of(GameN& g, uint8_t* field, uint8_t* vec, uint8_t* pidx) {
        Board<G, 1> bd[MAX_PLAYERS];
        bind(g, bd);
        for (int pl = 0; pl < np; pl++) {
            if (field)
                for (int c = 0; c < H * W; c++) field[pl * H * W + c] = (bd[pl].sw(c >> 5) >> (c & 31)) & 1;
            if (vec) observe_vec(g.p[pl], vec + pl * VEC_LEN);
            if (pidx) pidx[pl] = (uint8_t)g.p[pl].kind;
            for (int y = 0; y < H; y++)       // the two board forms (columns vs stream) must agree
                for (int x = 0; x < W; x++)
                    if ((int)((bd[pl].col(x) >> (y + G::YB)) & 1) != (int)((bd[pl].sw((y * W + x) >> 5) >> ((y * W + x) & 31)) & 1)) abort();
        }
    }
    void observe(const int32_t* idx, int cnt, uint8_t* field, uint8_t* vec, uint8_t* pidx) override {
        for (int i = 0; i < cnt; i++)
            obs_of(games[idx ? idx[i] : i], field ? field + (size_t)i * np * H * W : nullptr, vec ? vec + (size_t)i * np * VEC_LEN : nullptr,
                   pidx ? pidx + (size_t)np * i : nullptr);
    }
    void scal(const int32_t* idx, int cnt, int32_t* out) override {
        for (int i = 0; i < cnt; i++) {
            GameN& g = games[idx ? idx[i] : i];
            for (int pl = 0; pl < np; pl++) {
                int32_t* s = out + ((size_t)i * np + pl) * 8;
                const PS& q = g.p[pl];
                s[0] = q.x; s[1] = q.y; s[2] = q.inc_count; s[3] = q.remaining; s[4] = q.combo_count; s[5] = q.next; s[6] = q.reward; s[7] = q.dead;
            }
        }
    }
    int enum_game(GameN& g, int player, int bar_null, uint32_t* lst, int cap, int* bad) {
        Board<G, 1> bd[MAX_PLAYERS];
        bind(g, bd);
        if (player >= np) player = np - 1;
        ListSink<KW> sink;
        const int off = bar_null ? 0 : 1;
        for (int i = 0; i < KW; i++) lst[i] = 0;
        sink.init(lst + off * KW, cap - off);
        const PS& me = g.p[player];
        if (me.kind <= 6) {
            uint8_t pbuf[EnumGeo<G>::PCAP];
            Enumerator<G, 1, KW, ListSink<KW>> e(bd[player], tb, me.kind, sink, pbuf);
            e.run(me.rot);
        }
        *bad = sink.bad;
        if (sink.count == 0) { for (int i = 0; i < KW; i++) lst[i] = 0; return 1; }
        return sink.count + off;
    }
    static void unpack_seq(const uint32_t* w, uint8_t* keys, uint8_t* len, int lmax) {
        const int L = packed_len<KW>(w);
        memset(keys, 0, lmax);
        for (int i = 0; i < L && i < lmax; i++) keys[i] = (w[i >> 3] >> (4 * (i & 7))) & 15;
        *len = (uint8_t)(L < lmax ? L : lmax);
    }
    int enumerate(const int32_t* idx, int cnt, const uint8_t* player, int bar_null, int kmax, int lmax, int32_t* K, uint8_t* keys,
                  uint8_t* key_len) override {
        std::vector<uint32_t> lst((size_t)(kmax + 1) * KW);
        int nbad = 0;
        for (int i = 0; i < cnt; i++) {
            int bad;
            const int k = enum_game(games[idx ? idx[i] : i], player[i], bar_null, lst.data(), kmax, &bad);
            K[i] = k; nbad += bad;
            for (int a = 0; a < kmax; a++) {
                uint8_t* dst = keys + ((size_t)i * kmax + a) * lmax;
                if (a < k) unpack_seq(&lst[(size_t)a * KW], dst, &key_len[(size_t)i * kmax + a], lmax);
                else { memset(dst, 0, lmax); key_len[(size_t)i * kmax + a] = 0; }
            }
        }
        return nbad;
    }
    int expand(const int32_t* idx, int cnt, const uint8_t* player, int bar_null, int finalize, int time_ms, int kmax, int lmax, int32_t* K,
               uint8_t* keys, uint8_t* key_len, uint8_t* field, uint8_t* vec, uint8_t* pidx, int8_t* reward, uint8_t* done,
               uint8_t* lines, uint8_t* dead) override {
        std::vector<uint32_t> lst((size_t)(kmax + 1) * KW);
        int nbad = 0;
        for (int i = 0; i < cnt; i++) {
            const GameN& base = games[idx ? idx[i] : i];
            GameN g = base;
            int bad;
            const int k = enum_game(g, player[i], bar_null, lst.data(), kmax, &bad);
            K[i] = k; nbad += bad;
            for (int a = 0; a < kmax; a++) {
                const size_t o = (size_t)i * kmax + a;
                uint8_t* dst = keys + o * lmax;
                if (a >= k) { memset(dst, 0, lmax); key_len[o] = 0; continue; }
                unpack_seq(&lst[(size_t)a * KW], dst, &key_len[o], lmax);
                g = base;
                int r;
                const bool d = perform(g, player[i], &lst[(size_t)a * KW], packed_len<KW>(&lst[(size_t)a * KW]), finalize != 0, time_ms, &r);
                if (reward) reward[o] = (int8_t)r;
                if (done) done[o] = d;
                for (int pl = 0; pl < np; pl++) {
                    if (lines) lines[o * np + pl] = (uint8_t)g.p[pl].reward;
                    if (dead) dead[o * np + pl] = (uint8_t)g.p[pl].dead;
                }
                obs_of(g, field ? field + o * np * H * W : nullptr, vec ? vec + o * np * VEC_LEN : nullptr, pidx ? pidx + np * o : nullptr);
            }
        }
        return nbad;
    }
    int flags(int gi) override { int f = 0; for (int pl = 0; pl < np; pl++) f |= games[gi].p[pl].flags; return f; }
    void get_game(int, b2t_game_raw*, uint8_t*) override { abort(); }          // b2t_game_raw describes two players
    void set_game(int, const b2t_game_raw*, const uint8_t*) override { abort(); }
    void copy_game(int dst, HSBase* from, int src) override { games[dst] = static_cast<HSN*>(from)->games[src]; ensure_row(games[dst].seed16); }
    void words_of(int, uint32_t*) override { abort(); }
    int nwords() override { return 0; }
    void enable_tiles() override {}
    uint32_t* planes_of(int) override { return nullptr; }
    int blob_bytes() override { return (int)sizeof(GameN); }
    void get_blob(int gi, void* out) override { memcpy(out, &games[gi], sizeof(GameN)); }
    void set_blob(int gi, const void* in) override { memcpy(&games[gi], in, sizeof(GameN)); ensure_row(games[gi].seed16); }
};

extern "C" {
void* hs_create(int n, int H, int W, const uint8_t* map, uint32_t piece_len, uint32_t hole_len) {
#define X(h, w) if (H == h && W == w) return (HSBase*)new HS<h, w>(n, map, piece_len, hole_len);
    B2T_FOR_EACH_GEOMETRY(X)
#undef X
    return nullptr;
}
void* hs_create_players(int n, int n_players, int H, int W, const uint8_t* map, uint32_t piece_len, uint32_t hole_len) {
    if (n_players == 2) return hs_create(n, H, W, map, piece_len, hole_len);
    if (n_players < 1 || n_players > MAX_PLAYERS) return nullptr;
#define X(h, w) if (H == h && W == w) return (HSBase*)new HSN<h, w>(n, n_players, map, piece_len, hole_len);
    B2T_FOR_EACH_GEOMETRY(X)
#undef X
    return nullptr;
}
void hs_destroy(void* h) { delete (HSBase*)h; }
void hs_reset(void* h, const int32_t* idx, int n, const uint32_t* seeds) { ((HSBase*)h)->reset(idx, n, seeds); }
void hs_step(void* h, const int32_t* idx, int n, const uint8_t* player, const uint8_t* keys, const uint8_t* key_len, int lmax, int time_ms,
             int finalize, int8_t* reward, uint8_t* done, uint8_t* lines) {
    ((HSBase*)h)->step(idx, n, player, keys, key_len, lmax, time_ms, finalize, reward, done, lines);
}
void hs_observe(void* h, const int32_t* idx, int n, uint8_t* field, uint8_t* vec, uint8_t* pidx) { ((HSBase*)h)->observe(idx, n, field, vec, pidx); }
void hs_scal(void* h, const int32_t* idx, int n, int32_t* out) { ((HSBase*)h)->scal(idx, n, out); }
int hs_enumerate(void* h, const int32_t* idx, int n, const uint8_t* player, int bar_null, int kmax, int lmax, int32_t* K, uint8_t* keys,
                 uint8_t* key_len) {
    return ((HSBase*)h)->enumerate(idx, n, player, bar_null, kmax, lmax, K, keys, key_len);
}
int hs_expand(void* h, const int32_t* idx, int n, const uint8_t* player, int bar_null, int finalize, int time_ms, int kmax, int lmax,
              int32_t* K, uint8_t* keys, uint8_t* key_len, uint8_t* field, uint8_t* vec, uint8_t* pidx, int8_t* reward, uint8_t* done,
              uint8_t* lines, uint8_t* dead) {
    return ((HSBase*)h)->expand(idx, n, player, bar_null, finalize, time_ms, kmax, lmax, K, keys, key_len, field, vec, pidx, reward, done, lines, dead);
}
int hs_flags(void* h, int game) { return ((HSBase*)h)->flags(game); }
void hs_get_game(void* h, int game, b2t_game_raw* out, uint8_t* field) { ((HSBase*)h)->get_game(game, out, field); }
void hs_set_game(void* h, int game, const b2t_game_raw* in, const uint8_t* field) { ((HSBase*)h)->set_game(game, in, field); }
void hs_enable_tiles(void* h) { ((HSBase*)h)->enable_tiles(); }
void hs_copy_game(void* to, int dst, void* from, int src) { ((HSBase*)to)->copy_game(dst, (HSBase*)from, src); }
int hs_blob_bytes(void* h) { return ((HSBase*)h)->blob_bytes(); }
void hs_get_blob(void* h, int game, void* out) { ((HSBase*)h)->get_blob(game, out); }
void hs_set_blob(void* h, int game, const void* in) { ((HSBase*)h)->set_blob(game, in); }
}
