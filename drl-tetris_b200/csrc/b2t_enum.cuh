// b200tetris — placement enumerator: the ordered key sequences of TestField::getMask(2)
// (/root/reference/environment/game_backend/source/TestField.cpp:64-415) on column bit-boards.
//
// For a start column x and rotation r ONE collision mask answers "does the piece fit at (x, y, r)" for every
// height y, so the top drop is a count-trailing-zeros and the under-overhang ("finesse") candidates are the
// runs of free heights below it; only those rare candidates run the backward slide / wall-kick search, which
// is restated step for step (including the places where the reference does NOT restore the probe piece
// between a failed left search and the right search — those are observable in the emitted paths).
//
// Sequences are built as 4-bit key tokens (0..10 fit) in KW 32-bit words, token i at nibble i; every
// sequence ends with key 7 and contains no zero token, so its length is implied.
#pragma once
#include "b2t_core.cuh"

#if defined(B2T_COUNTERS) && !defined(__CUDA_ARCH__)
extern "C" long long b2t_counters[16];   // test-only instrumentation (tests/hostsim): 0 boards 1 candidates 2 reachable 3 fits in search
#define B2T_COUNT(i) (++b2t_counters[i])
#else
#define B2T_COUNT(i) do { } while (0)
#endif

namespace b2t {

template <int KW> struct Seq {
    uint32_t w[KW];
    int n;
    B2T_HD void clear() {
#pragma unroll
        for (int i = 0; i < KW; i++) w[i] = 0;
        n = 0;
    }
    B2T_HD void push(uint32_t tok) {  // append at nibble n
        const uint32_t v = tok << (4 * (n & 7));
#pragma unroll
        for (int i = 0; i < KW; i++) w[i] |= ((n >> 3) == i) ? v : 0u;   // select form: keeps w[] in registers
        n++;
    }
    B2T_HD void push_front(uint32_t tok) {  // prepend: everything moves up one nibble
#pragma unroll
        for (int i = KW - 1; i > 0; i--) w[i] = (w[i] << 4) | (w[i - 1] >> 28);
        w[0] = (w[0] << 4) | tok;
        n++;
    }
    B2T_HD uint32_t pop_front() {
        const uint32_t t = w[0] & 15u;
#pragma unroll
        for (int i = 0; i + 1 < KW; i++) w[i] = (w[i] >> 4) | (w[i + 1] << 28);
        w[KW - 1] >>= 4;
        n--;
        return t;
    }
};

// [rotation][|x - spawn| laterals][7] as one 64-bit token string (<= 16 tokens: W <= 32), TestField.cpp:113-121
B2T_HD uint64_t top_drop_tokens(int rot_delta, int dx) {
    const int an = dx < 0 ? -dx : dx;
    uint64_t s = (dx > 0 ? 0x3333333333333333ull : 0x1111111111111111ull) & ((1ull << (4 * an)) - 1ull);
    s |= 7ull << (4 * an);
    if (rot_delta) s = (s << 4) | (uint64_t)(rot_delta == 1 ? 8u : rot_delta == 2 ? 10u : 9u);
    return s;
}

template <int KW> B2T_HD bool seq_equal(const Seq<KW>& a, const uint32_t* b) {
    bool eq = true;
#pragma unroll
    for (int i = 0; i < KW; i++) eq = eq && (a.w[i] == b[i]);
    return eq;
}

// Token writers: put(key) appends one 4-bit key token.
template <int KW> struct SeqWriter {      // into a register Seq (host paths)
    Seq<KW>& s;
    bool bad;
    B2T_HD explicit SeqWriter(Seq<KW>& s_) : s(s_), bad(false) { s.clear(); }
    B2T_HD void put(uint32_t tok) { if (s.n < KW * 8) s.push(tok); else bad = true; }
    B2T_HD void put_run(uint32_t tok, int count) { for (int i = 0; i < count; i++) put(tok); }
};
template <int KW, int WS = 1> struct WordWriter {     // into KW words (stride WS) in (shared) memory: dynamic word index costs nothing there
    uint32_t* w;
    uint32_t acc;
    int n;
    bool bad;
    B2T_HD explicit WordWriter(uint32_t* w_) : w(w_), acc(0u), n(0), bad(false) {}
    B2T_HD void put(uint32_t tok) {
        if (n >= KW * 8) { bad = true; return; }
        acc |= tok << (4 * (n & 7));
        n++;
        if ((n & 7) == 0) { w[((n >> 3) - 1) * WS] = acc; acc = 0u; }
    }
    B2T_HD void put_run(uint32_t tok, int count) {   // `count` copies of one token, a word at a time
        while (count > 0) {
            if (n >= KW * 8) { bad = true; return; }
            const int at = n & 7, room = 8 - at;
            const int m = count < room ? count : room;
            const uint32_t mask = (m == 8) ? 0xFFFFFFFFu : ((1u << (4 * m)) - 1u);
            acc |= ((tok * 0x11111111u) & mask) << (4 * at);
            n += m;
            count -= m;
            if ((n & 7) == 0) { w[((n >> 3) - 1) * WS] = acc; acc = 0u; }
        }
    }
    B2T_HD void finish() {   // flush the partial word and zero the rest
        int j = n >> 3;
        if (n & 7) w[(j++) * WS] = acc;
        for (; j < KW; j++) w[j * WS] = 0u;
    }
};

// Sink concept:  void top(const Seq<KW>&, bool bad);  void finesse(const Seq<KW>&, bool bad);
// PST = stride of the per-thread path buffer (1 on the host, block size in shared memory)
template <class G> struct EnumGeo { static constexpr int PCAP = 2 * G::W + 6; };   // path entries: turn + rise + (lateral + rise) per column

template <class G, int S, int KW, class Sink, int PST = 1> struct Enumerator {
    typedef typename G::col_t col_t;
    Board<G, S> b;      // by value (two pointers): the kernel re-points it at the owner of each pooled candidate
    Sink& sink;
    int kind, spawn_rot;
    uint64_t pcn4, pcc4;   // the four rotations of this piece, 16 bits each
    int x, y, rot;         // the probe piece of TestField
    // test_path (TestField.h:36), one byte per entry like the reference, as FORWARD keys: 3 / 1 = the lateral move undone by
    // a left / right search step (255 / 254, :6-12), 6 = 253, 8/10/9 = 241/242/243, 16+n = "n rows up" (n times key 5)
    uint8_t* pbuf;
    int path_n;
    bool path_bad;
    int best_x, best_rot;  // finesseMove (:166-172); its path is the test_path at the moment of success (the search stops there)

    B2T_HD Enumerator(const Board<G, S>& b_, const Tabs& tb, int kind_, Sink& sink_, uint8_t* pbuf_) : b(b_), sink(sink_), kind(kind_), pbuf(pbuf_) {
        spawn_rot = spawn_rot_of(tb, kind);
        pcn4 = pcc4 = 0;
        for (int r = 0; r < 4; r++) {
            pcn4 |= (uint64_t)tb.pcn[kind * 4 + r] << (16 * r);
            pcc4 |= (uint64_t)tb.pcc[kind * 4 + r] << (16 * r);
        }
        path_n = 0; path_bad = false;
        best_x = best_rot = 0; x = y = rot = 0;
    }
    // light re-targeting for pooled searches: another board / piece, tables already packed by the owner
    B2T_HD void rebind(const Board<G, S>& b_, int kind_, int spawn_rot_, uint64_t pcn4_, uint64_t pcc4_) {
        b = b_; kind = kind_; spawn_rot = spawn_rot_; pcn4 = pcn4_; pcc4 = pcc4_;
    }
    B2T_HD uint32_t pcn(int r) const { return (uint32_t)(pcn4 >> (16 * r)) & 0xFFFFu; }
    B2T_HD uint32_t pcc(int r) const { return (uint32_t)(pcc4 >> (16 * r)) & 0xFFFFu; }
    B2T_HD bool fit_at(int xx, int yy, int rr) const { B2T_COUNT(3); return fits<G, S>(b, xx, yy, pcn(rr)); }
    B2T_HD bool fit() const { return fit_at(x, y, rot); }

    // ---- test_path bookkeeping
    B2T_HD int path_entries() const { return path_n; }
    B2T_HD void path_clear() { path_n = 0; path_bad = false; }
    B2T_HD void path_entry(uint32_t e) {
        if (path_n >= EnumGeo<G>::PCAP) { path_bad = true; return; }
        pbuf[path_n * PST] = (uint8_t)e;
        path_n++;
    }
    B2T_HD void path_lateral(bool searching_left) { path_entry(searching_left ? 3u : 1u); }
    B2T_HD void path_turn(int turn) { path_entry(turn == 1 ? 8u : turn == 2 ? 10u : 9u); }
    B2T_HD void path_up(int up) {
        if (!up) return;
        path_entry(up == 253 ? 6u : 16u + (uint32_t)up);
    }
    B2T_HD bool record() { best_x = x; best_rot = rot; return true; }  // setFinesseMove

    // moveUp (:392-410): rise while the piece fits; 253 if it rose and was resting on something, else rows risen
    B2T_HD int move_up() {
        const bool blocked_below = !fit_at(x, y + 1, rot);
        const col_t cm = collide_mask<G, S>(b, x, pcc(rot));
        const col_t above = (col_t)(cm & ((G::one() << (y + G::YB)) - 1));   // collisions strictly above y
        const int ny = msb(above) - G::YB + 1;                                // ceiling rows always collide
        const int count = y - ny;
        y = ny;
        if (count && blocked_below) return 253;
        return count;
    }
    // ---------------------------------------------------------------------------------------------------------------
    // finesseIsPossible (:189-200) and everything below it (reverseWallkick :261-278, doWallKick :280-356, r180KeepPos
    // :358-390, tryUp :240-259, tryLeft/tryRight :202-238) as ONE state machine, so that 32 lanes searching 32 different
    // cavities share every instruction of the expensive primitives: search_step() runs, in this order, at most one
    // "8 poses around a base" test (the inverse wall-kick probe or the forward-rotation check) and at most one slide
    // (lateral moves with the rise after each), whatever stage each lane is in.  The observable quirks are kept: the right
    // slide starts where the failed left slide stopped (:196-199, :252-258), paths are cleared only where the reference
    // clears them, an in-place fit of a rotated pose ends doWallKick whatever tryUp returns (:294-302).
    enum { ST_SIDE, ST_WK_R, ST_FWD, ST_TU, ST_TU_R, ST_WK_L, ST_WK_RR, ST_WK_FAIL, ST_FINAL_L, ST_FINAL_R, ST_DONE };
    int st, sx, sy, sr, bx, by, br, px, py, pr, rp, turn, up, attempt, inplace;
    bool found;

    B2T_HD void search_begin(int cx, int cy, int cr) {
        sx = x = cx; sy = y = cy; sr = rot = cr;
        st = (kind == 6) ? ST_FINAL_L : ST_SIDE;
        found = false;
        bx = by = br = px = py = pr = rp = turn = up = attempt = inplace = 0;
    }
    // bit 0: fits at (bx0, by0, r); bit k+1: fits at the k-th kick offset.  The 7 offsets are those of kickTest
    // (gameField.cpp:93-103), cumulative (dy,dx) = (1,0)(0,-1)(0,1)(1,-1)(1,1)(0,-2)(0,2); `inverse` negates dy (:304-310)
    B2T_HD uint32_t fit8(int bx0, int by0, int r, bool inverse) const {
        typedef typename G::col_t col_t;
        B2T_COUNT(4);
        B2T_CHECK_XY(bx0 - 2, by0 - 1); B2T_CHECK_XY(bx0 + 2, by0 + 1);
        const uint32_t n = pcn(r);
        const col_t* p = b.cols + (bx0 - 2 + G::XL) * S;
        col_t c[8];
#pragma unroll
        for (int j = 0; j < 8; j++) c[j] = p[j * S];
        const int s0 = by0 + G::YB, s1 = inverse ? s0 - 1 : s0 + 1;
        uint32_t m = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int dx = (k == 0) ? 0 : (k == 2 || k == 4) ? -1 : (k == 3 || k == 5) ? 1 : (k == 6) ? -2 : (k == 7) ? 2 : 0;
            const bool dyk = (k == 1 || k == 4 || k == 5);
            const int sh = dyk ? s1 : s0;
            col_t acc = (c[dx + 2] >> sh) & (col_t)(n & 15u);
            acc |= (c[dx + 3] >> sh) & (col_t)((n >> 4) & 15u);
            acc |= (c[dx + 4] >> sh) & (col_t)((n >> 8) & 15u);
            acc |= (c[dx + 5] >> sh) & (col_t)((n >> 12) & 15u);
            if (acc == 0) m |= 1u << k;
        }
        return m;
    }
    B2T_HD static int kick_dx(int k) { return (int)((0x1032CAu >> (3 * k)) & 7u) - 2; }   // k = 0..6
    B2T_HD static int kick_dy(int k) { return (0x19 >> k) & 1; }
    B2T_HD void flip_pose(int& fx, int& fy, int& fr) const {  // r180KeepPos, :358-390 (I,Z shift one way, S the other)
        fr = (fr + 2) & 3;
        const int s = (kind == 2) ? -1 : 1;
        if (fr == 0) fx += s;
        else if (fr == 1) fy += s;
        else if (fr == 2) fx -= s;
        else fy -= s;
    }
    // one iteration; true when the search is over (result in `found`; on success path/x/rot hold the recorded move)
    B2T_HD bool search_step() {
        B2T_COUNT(5);
        // ---- cheap stage transitions
        if (st == ST_SIDE) {                                   // reverseWallkick :262-266
            const bool fr = fit_at(sx + 1, sy, sr), fl = fit_at(sx - 1, sy, sr);
            if (fr || fl) st = ST_FINAL_L;
            else { attempt = 0; bx = sx; by = sy; br = sr; rp = 0; st = ST_WK_R; }
        }
        if (st == ST_WK_FAIL) {
            if (attempt == 0 && (kind == 2 || kind == 3 || kind == 4)) {   // :272-275
                attempt = 1; bx = sx; by = sy; br = sr;
                flip_pose(bx, by, br);
                rp = 0; st = ST_WK_R;
            } else st = ST_FINAL_L;
        }
        // ---- the 8-pose primitive
        bool want8 = false, inv = true;
        int fx = 0, fy = 0, fr8 = 0;
        if (st == ST_WK_R) {
            if (rp == br) rp++;
            if (rp > 3) st = ST_WK_FAIL;
            else { want8 = true; fx = bx; fy = by; fr8 = rp; inv = true; }
        } else if (st == ST_FWD) { want8 = true; fx = px; fy = py; fr8 = br; inv = false; }
        if (want8) {
            const uint32_t m8 = fit8(fx, fy, fr8, inv);
            if (st == ST_WK_R) {
                turn = (br - rp) & 3;
                if (m8 & 1u) { x = bx; y = by; rot = rp; inplace = 1; st = ST_TU; }                    // :294-302
                else if (m8 >> 1) {                                                                     // :304-310
                    const int k = ctz32(m8 >> 1);
                    px = bx + kick_dx(k); py = by - kick_dy(k); pr = rp;
                    st = ST_FWD;
                    return false;
                } else { rp++; return false; }
            } else {                                                                                    // :323-333
                bool good = m8 != 0;
                if (good) {
                    const int k = ctz32(m8);
                    const int lx = px + (k ? kick_dx(k - 1) : 0), ly = py + (k ? kick_dy(k - 1) : 0);
                    good = (lx == bx && ly == by);
                }
                if (!good) { st = ST_WK_FAIL; return false; }
                x = px; y = py; rot = pr; inplace = 0; st = ST_TU;
            }
        }
        // ---- the slide primitive
        int dir = 0, on_fail = ST_DONE;
        bool clear = false, rise_first = false;
        if (st == ST_TU) {                  // tryUp :240-252
            path_clear(); path_turn(turn);
            dir = -1; rise_first = true; on_fail = ST_TU_R;
        } else if (st == ST_TU_R) {         // tryUp :254-258 (from wherever the left slide stopped)
            path_clear(); path_turn(turn); path_up(up);
            dir = 1; on_fail = inplace ? ST_WK_FAIL : ST_WK_L;
        } else if (st == ST_WK_L) {         // doWallKick :340-344
            x = px; y = py; rot = pr;
            path_clear(); path_turn(turn);
            dir = -1; on_fail = ST_WK_RR;
        } else if (st == ST_WK_RR) {        // doWallKick :346-352
            x = px; y = py; rot = pr;
            if (path_entries() > 1) { path_clear(); path_turn(turn); }
            dir = 1; on_fail = ST_WK_FAIL;
        } else if (st == ST_FINAL_L) {      // finesseIsPossible :195-196
            x = sx; y = sy; rot = sr;
            dir = -1; clear = true; on_fail = ST_FINAL_R;
        } else if (st == ST_FINAL_R) {      // finesseIsPossible :199 (not restored)
            dir = 1; clear = true; on_fail = ST_DONE;
        }
        if (dir != 0) {
            B2T_COUNT(6);
            bool success = false;
            for (;;) {
                B2T_COUNT(7);
                if (!rise_first) {
                    if (!fit_at(x + dir, y, rot)) break;
                    x += dir;
                    if (clear) { path_clear(); clear = false; }
                    path_lateral(dir < 0);
                }
                const int u = move_up();
                if (rise_first) up = u;
                path_up(u);
                rise_first = false;
                if (y == 0) { success = true; break; }
            }
            if (success) { record(); found = true; st = ST_DONE; }
            else st = on_fail;
        }
        return st == ST_DONE;
    }
    B2T_HD bool reachable() {
        search_begin(x, y, rot);
        while (!search_step()) {}
        x = sx; y = sy; rot = sr;
        return found;
    }

    // ---- emission (:37-49, :113-164)
    B2T_HD void emit_top(int r, int xx) {
        Seq<KW> s; s.clear();
        const int dx = xx - G::SPAWN_X;
        const uint64_t t = top_drop_tokens((r - spawn_rot) & 3, dx);
        s.w[0] = (uint32_t)t; s.w[1] = (uint32_t)(t >> 32);
        s.n = (dx < 0 ? -dx : dx) + 1 + (((r - spawn_rot) & 3) ? 1 : 0);
        sink.top(s, false);
    }
    // forward key string of the recorded finesse move (useFinesseMove, :127-164): rotation, laterals to the column the
    // search ended in, the path walked backwards (:152-158), hard drop
    template <class Wr> B2T_HD void write_finesse(Wr& w) const {
        write_finesse_raw<Wr, PST>(w, pbuf, path_n, best_x, best_rot, spawn_rot);
    }
    template <class Wr, int PS2>
    B2T_HD static void write_finesse_raw(Wr& w, const uint8_t* pb, int pn, int bx_, int brot, int spawn) {
        const int d = (brot - spawn) & 3;
        if (d) w.put(d == 1 ? 8u : d == 2 ? 10u : 9u);
        const int n = bx_ - G::SPAWN_X;
        w.put_run(n > 0 ? 3u : 1u, n > 0 ? n : -n);
        for (int i = pn - 1; i >= 0; i--) {
            const uint32_t e = pb[i * PS2];
            if (e >= 16u) w.put_run(5u, (int)e - 16);
            else w.put(e);
        }
        w.put(7u);
    }
    B2T_HD void emit_finesse() {
        Seq<KW> s;
        SeqWriter<KW> w(s);
        write_finesse(w);
        sink.finesse(s, w.bad || path_bad);
    }

    // getMask (:64-111) + findNextMove (:113-125) + tryAllFinesseMoves (:174-187)
    B2T_HD void run(int cur_rot) {
        B2T_COUNT(0);
        const int nrot = (kind == 6) ? 1 : (kind == 4 || kind == 2 || kind == 3) ? 2 : 4;
        for (int xx = -1; xx < G::W - 1; xx++)
            for (int r0 = 0; r0 < nrot; r0++) {
                const int r = (kind == 6) ? cur_rot : r0;
                const col_t cm = collide_mask<G, S>(b, xx, pcc(r));
                if ((cm >> G::YB) & 1) continue;
                emit_top(r, xx);
                const int y0 = rest_from<G>(cm, 0);
                int yy = y0 + 2;
                while (yy < G::H - 1) {
                    col_t free_ = (col_t)(~cm >> (yy + G::YB));
                    free_ &= (col_t)((G::one() << (G::H - 1 - yy)) - 1);
                    if (!free_) break;
                    const int a = yy + ctz(free_);
                    const int rest = rest_from<G>(cm, a);
                    x = xx; y = rest; rot = r;
                    B2T_COUNT(1);
                    if (reachable()) { B2T_COUNT(2); emit_finesse(); }
                    yy = rest + 1;
                }
            }
    }
};

}  // namespace b2t
