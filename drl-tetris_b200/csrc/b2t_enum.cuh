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

// Sink concept:  void top(const Seq<KW>&);  void finesse(const Seq<KW>&);   (overflowing sequences set .bad)
template <class G, int S, int KW, class Sink> struct Enumerator {
    typedef typename G::col_t col_t;
    const Board<G, S>& b;
    Sink& sink;
    int kind, spawn_rot;
    uint64_t pcn4, pcc4;   // the four rotations of this piece, 16 bits each
    int x, y, rot;         // the probe piece of TestField
    // test_path: the FORWARD keys of the path searched so far, newest move first (= reversed path, :152-158)
    Seq<KW> path;
    int path_entries;      // test_path.size() of the reference (an "n rows up" entry counts once)
    bool path_bad;
    // finesseMove (:166-172)
    Seq<KW> best;
    int best_x, best_rot;
    bool best_bad;

    B2T_HD Enumerator(const Board<G, S>& b_, const Tabs& tb, int kind_, Sink& sink_) : b(b_), sink(sink_), kind(kind_) {
        spawn_rot = spawn_rot_of(tb, kind);
        pcn4 = pcc4 = 0;
        for (int r = 0; r < 4; r++) {
            pcn4 |= (uint64_t)tb.pcn[kind * 4 + r] << (16 * r);
            pcc4 |= (uint64_t)tb.pcc[kind * 4 + r] << (16 * r);
        }
        path.clear(); best.clear();
        path_entries = 0; path_bad = best_bad = false;
        best_x = best_rot = 0; x = y = rot = 0;
    }
    B2T_HD uint32_t pcn(int r) const { return (uint32_t)(pcn4 >> (16 * r)) & 0xFFFFu; }
    B2T_HD uint32_t pcc(int r) const { return (uint32_t)(pcc4 >> (16 * r)) & 0xFFFFu; }
    B2T_HD bool fit_at(int xx, int yy, int rr) const { B2T_COUNT(3); return fits<G, S>(b, xx, yy, pcn(rr)); }
    B2T_HD bool fit() const { return fit_at(x, y, rot); }

    // ---- test_path bookkeeping
    B2T_HD void path_clear() { path.clear(); path_entries = 0; path_bad = false; }
    B2T_HD void path_key(uint32_t key) {
        if (path.n >= KW * 8) { path_bad = true; return; }
        path.push_front(key);
    }
    B2T_HD void path_lateral(bool searching_left) { path_key(searching_left ? 3u : 1u); path_entries++; }  // 255->3, 254->1 (:6-12)
    B2T_HD void path_turn(int turn) { path_key(turn == 1 ? 8u : turn == 2 ? 10u : 9u); path_entries++; }     // 241/242/243 (:22-32)
    B2T_HD void path_up(int up) {  // 253 -> key 6, n -> n times key 5 (:14-20, :153-156)
        if (!up) return;
        if (up == 253) path_key(6u);
        else for (int i = 0; i < up; i++) path_key(5u);
        path_entries++;
    }
    B2T_HD bool record() { best = path; best_bad = path_bad; best_x = x; best_rot = rot; return true; }  // setFinesseMove

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
    B2T_HD bool try_left(bool clear) {  // :202-220
        for (;;) {
            if (!fit_at(x - 1, y, rot)) return false;
            x--;
            if (clear) path_clear();
            path_lateral(true);
            path_up(move_up());
            if (y == 0) return record();
            clear = false;
        }
    }
    B2T_HD bool try_right(bool clear) {  // :222-238
        for (;;) {
            if (!fit_at(x + 1, y, rot)) return false;
            x++;
            if (clear) path_clear();
            path_lateral(false);
            path_up(move_up());
            if (y == 0) return record();
            clear = false;
        }
    }
    B2T_HD bool try_up(int turn) {  // :240-259
        path_clear();
        const int up = move_up();
        path_turn(turn);
        path_up(up);
        if (y == 0) return record();
        if (try_left(false)) return true;
        path_clear();
        path_turn(turn);
        path_up(up);
        return try_right(false);
    }
    // BasicField::rcw/rccw/r180 on the probe piece (gameField.cpp:55-103)
    B2T_HD bool forward_rotate(int turn) {
        const int nr = (rot + turn) & 3;
        if (fit_at(x, y, nr)) { rot = nr; return true; }
#pragma unroll
        for (int k = 0; k < 7; k++) {
            const int dy = (k == 0 || k == 3 || k == 4) ? 1 : 0;
            const int dx = (k == 1 || k == 3) ? -1 : (k == 2 || k == 4) ? 1 : (k == 5) ? -2 : (k == 6) ? 2 : 0;
            if (fit_at(x + dx, y + dy, nr)) { x += dx; y += dy; rot = nr; return true; }
        }
        return false;
    }
    B2T_HD bool wall_kick() {  // doWallKick, :280-356
        const int sx = x, sy = y, sr = rot;
        int r;
        bool working = false;
        for (r = 0; r < 4; r++) {
            working = true;
            if (r == sr) { working = false; continue; }
            rot = r;
            if (fit()) {
                const bool ok = try_up((sr - r) & 3);
                x = sx; y = sy; rot = sr;
                return ok;
            }
            // inverse kick offsets, cumulative (dy,dx): (-1,0)(0,-1)(0,1)(-1,-1)(-1,1)(0,-2)(0,2)
            bool found = false;
#pragma unroll
            for (int k = 0; k < 7; k++) {
                const int dy = (k == 0 || k == 3 || k == 4) ? -1 : 0;
                const int dx = (k == 1 || k == 3) ? -1 : (k == 2 || k == 4) ? 1 : (k == 5) ? -2 : (k == 6) ? 2 : 0;
                if (!found && fit_at(sx + dx, sy + dy, r)) { found = true; x = sx + dx; y = sy + dy; }
            }
            if (found) break;
            x = sx; y = sy;
            working = false;
        }
        if (!working) { x = sx; y = sy; rot = sr; return false; }
        const int turn = (sr - r) & 3;
        const int px = x, py = y, pr = rot;   // candidate pre-pose
        forward_rotate(turn);
        if (x != sx || y != sy) { x = sx; y = sy; rot = sr; return false; }
        x = px; y = py; rot = pr;
        if (try_up(turn)) { x = sx; y = sy; rot = sr; return true; }
        x = px; y = py; rot = pr;
        path_clear();
        path_turn(turn);
        if (try_left(false)) { x = sx; y = sy; rot = sr; return true; }
        x = px; y = py; rot = pr;
        if (path_entries > 1) { path_clear(); path_turn(turn); }
        const bool ok = try_right(false);
        x = sx; y = sy; rot = sr;
        return ok;
    }
    B2T_HD void flip_in_place() {  // r180KeepPos, :358-390 (I,Z shift one way, S the other)
        rot = (rot + 2) & 3;
        const int s = (kind == 2) ? -1 : 1;
        if (rot == 0) x += s;
        else if (rot == 1) y += s;
        else if (rot == 2) x -= s;
        else y -= s;
    }
    B2T_HD bool reverse_kick() {  // reverseWallkick, :261-278
        if (kind == 6) return false;
        if (fit_at(x + 1, y, rot)) return false;
        if (fit_at(x - 1, y, rot)) return false;
        const int sx = x, sy = y, sr = rot;
        if (wall_kick()) { x = sx; y = sy; rot = sr; return true; }
        bool ok = false;
        if (kind == 2 || kind == 3 || kind == 4) {
            flip_in_place();
            ok = wall_kick();
        }
        x = sx; y = sy; rot = sr;
        return ok;
    }
    B2T_HD bool reachable() {  // finesseIsPossible, :189-200
        const int sx = x, sy = y, sr = rot;
        bool ok = reverse_kick();
        if (!ok) {
            x = sx; y = sy; rot = sr;
            ok = try_left(true);
            if (!ok) ok = try_right(true);   // NOT restored in between (:196-199)
        }
        x = sx; y = sy; rot = sr;
        return ok;
    }

    // ---- emission (:37-49, :113-164)
    B2T_HD void emit_prefix(Seq<KW>& s, int r, int xx) const {
        const int d = (r - spawn_rot) & 3;
        if (d) s.push(d == 1 ? 8u : d == 2 ? 10u : 9u);
        const int n = xx - G::SPAWN_X;
        for (int i = 0; i < (n > 0 ? n : -n); i++) s.push(n > 0 ? 3u : 1u);
    }
    B2T_HD void emit_top(int r, int xx) {
        Seq<KW> s; s.clear();
        const int dx = xx - G::SPAWN_X;
        const uint64_t t = top_drop_tokens((r - spawn_rot) & 3, dx);
        s.w[0] = (uint32_t)t; s.w[1] = (uint32_t)(t >> 32);
        s.n = (dx < 0 ? -dx : dx) + 1 + (((r - spawn_rot) & 3) ? 1 : 0);
        sink.top(s, false);
    }
    // forward key sequence of the recorded finesse move (useFinesseMove, :127-164); returns true if it did not fit
    B2T_HD bool build_finesse(Seq<KW>& s) const {
        s.clear();
        emit_prefix(s, best_rot, best_x);
        Seq<KW> p = best;
        const bool bad = best_bad || (s.n + p.n + 1 > KW * 8);
        while (p.n > 0 && s.n < KW * 8 - 1) s.push(p.pop_front());
        s.push(7u);
        return bad;
    }
    B2T_HD void emit_finesse() {
        Seq<KW> s;
        const bool bad = build_finesse(s);
        sink.finesse(s, bad);
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
