/* TEST INFRASTRUCTURE ONLY — see tetris_oracle.h.  Plain C11, no dependencies beyond libm.
 *
 * Every function cites the reference file:line (relative to /root/reference/) it restates.
 * SRC = environment/game_backend/source.
 */
#include "tetris_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ MT19937 (std::mt19937) */
typedef struct {
    uint32_t s[624];
    int at;
} mt_t;

static void mt_seed(mt_t *m, uint32_t seed) {
    m->s[0] = seed;
    for (int i = 1; i < 624; i++) m->s[i] = 1812433253u * (m->s[i - 1] ^ (m->s[i - 1] >> 30)) + (uint32_t)i;
    m->at = 624;
}

static uint32_t mt_next(mt_t *m) {
    if (m->at >= 624) {
        for (int i = 0; i < 624; i++) {
            uint32_t y = (m->s[i] & 0x80000000u) | (m->s[(i + 1) % 624] & 0x7fffffffu);
            m->s[i] = m->s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        m->at = 0;
    }
    uint32_t y = m->s[m->at++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* SRC/randomizer.h:8-34 UniformRealDistribution<float>(0,1): double scale 2^-32, result narrowed to float */
static float mt_unit_float(mt_t *m) {
    double d = (double)mt_next(m) * (1.0 / 4294967296.0) + 0.0;
    return (float)d;
}

/* ------------------------------------------------------------------ types */
typedef struct {
    uint8_t grid[4][4];
    uint8_t rotation;         /* spawn rotation */
    uint8_t current_rotation;
    int8_t x, y;
    uint8_t big;              /* lpiece: rotates as 4x4 */
    uint8_t tile, kind;
} piece_t;

typedef struct {
    int16_t count;
    int32_t delay;
} garb_t;

#define GARB_CAP 256
#define PATH_CAP 512

typedef struct {
    /* BasicField (SRC/gameField.h:16-45) */
    uint8_t *cell; /* [H*W] */
    piece_t cur;
    /* randomizer (SRC/randomizer.h:36-63) */
    float weight[7];
    mt_t hole_gen, piece_gen;
    /* GameplayData: only linesCleared is observable (through reward) */
    uint16_t data_lines_cleared;
    /* GarbageHandler (SRC/Garbage.h:15-37) */
    garb_t q[GARB_CAP];
    int qn;
    int32_t min_remaining, add_delay;
    /* ComboCounter (SRC/Combo.h:8-33) */
    int32_t combo_start, combo_time;
    uint8_t line_count, combo_count;
    uint16_t remaining;
    /* DropDelay (SRC/DropDelay.h:6-18) */
    int32_t drop_delay, drop_delay_time, inc_drop_delay, inc_drop_delay_time, lockdown_time;
    uint8_t lockdown;
    /* GamePlay scalars (SRC/gamePlay.h:58-69) */
    uint8_t nextpiece;
    float incoming_lines;
    uint8_t incoming_lines_count;
    int32_t time_ms;
    uint16_t lines_cleared;
    uint8_t reward, dead;
} player_t;

typedef struct {
    player_t pl[TOR_MAX_PLAYERS];
    uint8_t round_over;
    int last_winner;
    uint8_t env_done; /* tetris_environment.done */
} game_t;

struct tor_vec {
    int n, np, H, W;
    uint8_t map[7];
    uint8_t only_zs;
    piece_t base[7];
    game_t *g;
    uint8_t *cells; /* n*np*H*W */
};

/* ------------------------------------------------------------------ pieces (SRC/pieces.cpp:5-53) */
static void piece_rcw(piece_t *p) {
    int n = p->big ? 4 : 3;
    uint8_t t[4][4];
    memcpy(t, p->grid, 16);
    for (int yy = 0; yy < n; yy++)
        for (int xx = 0; xx < n; xx++) p->grid[xx][n - 1 - yy] = t[yy][xx];
    p->current_rotation = (uint8_t)((p->current_rotation + 1) & 3);
}

static void piece_rccw(piece_t *p) {
    int n = p->big ? 4 : 3;
    uint8_t t[4][4];
    memcpy(t, p->grid, 16);
    for (int yy = 0; yy < n; yy++)
        for (int xx = 0; xx < n; xx++) p->grid[n - 1 - xx][yy] = t[yy][xx];
    p->current_rotation = (uint8_t)((p->current_rotation + 3) & 3);
}

/* SRC/gamePlay.cpp:116-158 initBasePieces + setPieceOrientation */
static void make_base_pieces(piece_t base[7]) {
    /* cells of the rotation-0 grids, as (row, col) lists; colour = grid value */
    static const uint8_t colour[7] = {4, 3, 5, 7, 2, 1, 6};
    static const uint8_t cells[7][4][2] = {
        {{0, 1}, {1, 1}, {2, 1}, {2, 2}}, /* L */
        {{0, 1}, {1, 1}, {2, 0}, {2, 1}}, /* J */
        {{0, 1}, {1, 1}, {1, 2}, {2, 2}}, /* S */
        {{0, 1}, {1, 0}, {1, 1}, {2, 0}}, /* Z */
        {{0, 1}, {1, 1}, {2, 1}, {3, 1}}, /* I */
        {{1, 0}, {1, 1}, {1, 2}, {2, 1}}, /* T */
        {{1, 1}, {1, 2}, {2, 1}, {2, 2}}, /* O */
    };
    static const uint8_t spawn_rot[7] = {3, 1, 3, 1, 1, 2, 0};
    for (int k = 0; k < 7; k++) {
        piece_t *p = &base[k];
        memset(p, 0, sizeof *p);
        for (int c = 0; c < 4; c++) p->grid[cells[k][c][0]][cells[k][c][1]] = colour[k];
        p->big = (k == 4 || k == 6);
        p->tile = (uint8_t)(k + 1);
        p->kind = (uint8_t)k;
        p->rotation = spawn_rot[k];
        while (p->current_rotation != p->rotation) piece_rcw(p);
    }
}

/* ------------------------------------------------------------------ BasicField (SRC/gameField.cpp) */
typedef struct {
    uint8_t *cell;
    piece_t *pc;
    int H, W;
} field_t;

/* gameField.cpp:10-20 */
static int fits(const field_t *f) {
    const piece_t *p = f->pc;
    for (int xx = 0; xx < 4; xx++)
        for (int yy = 0; yy < 4; yy++)
            if (p->grid[yy][xx]) {
                int cx = p->x + xx, cy = p->y + yy;
                if (cx < 0 || cx > f->W - 1 || cy < 0 || cy > f->H - 1) return 0;
                if (f->cell[cy * f->W + cx]) return 0;
            }
    return 1;
}
/* gameField.cpp:22-47 */
static int move_right(field_t *f) { f->pc->x++; if (fits(f)) return 1; f->pc->x--; return 0; }
static int move_left(field_t *f) { f->pc->x--; if (fits(f)) return 1; f->pc->x++; return 0; }
static int move_down(field_t *f) { f->pc->y++; if (fits(f)) return 1; f->pc->y--; return 0; }
/* gameField.cpp:49-53 */
static void drop(field_t *f) { do { f->pc->y++; } while (fits(f)); f->pc->y--; }
/* gameField.cpp:93-103 */
static int kick_test(field_t *f) {
    piece_t *p = f->pc;
    p->y++; if (fits(f)) return 1;
    p->y--; p->x--; if (fits(f)) return 1;
    p->x += 2; if (fits(f)) return 1;
    p->y++; p->x -= 2; if (fits(f)) return 1;
    p->x += 2; if (fits(f)) return 1;
    p->x -= 3; p->y--; if (fits(f)) return 1;
    p->x += 4; if (fits(f)) return 1;
    return 0;
}
/* gameField.cpp:55-91 */
static int rot_cw(field_t *f) {
    piece_rcw(f->pc);
    if (fits(f) || kick_test(f)) return 1;
    f->pc->x -= 2; piece_rccw(f->pc); return 0;
}
static int rot_ccw(field_t *f) {
    piece_rccw(f->pc);
    if (fits(f) || kick_test(f)) return 1;
    f->pc->x -= 2; piece_rcw(f->pc); return 0;
}
static int rot_180(field_t *f) {
    piece_rccw(f->pc); piece_rccw(f->pc);
    if (fits(f) || kick_test(f)) return 1;
    f->pc->x -= 2; piece_rcw(f->pc); piece_rcw(f->pc); return 0;
}
/* gameField.cpp:105-110 */
static void stamp(field_t *f) {
    const piece_t *p = f->pc;
    for (int xx = 0; xx < 4; xx++)
        for (int yy = 0; yy < 4; yy++)
            if (p->grid[yy][xx]) f->cell[(p->y + yy) * f->W + p->x + xx] = p->tile;
}
/* gameField.cpp:112-118 (the read of row -1 at :115 has no visible effect: row 0 is zeroed at :116-117) */
static void remove_line(field_t *f, int row) {
    for (int yy = row; yy > 0; yy--) memcpy(f->cell + yy * f->W, f->cell + (yy - 1) * f->W, (size_t)f->W);
    memset(f->cell, 0, (size_t)f->W);
}
/* gameField.cpp:120-145; returns lines, *garbage_lines = rows containing an 8 */
static int clear_lines(field_t *f, int *garbage_lines) {
    int lines = 0;
    *garbage_lines = 0;
    for (int yy = f->H - 1; yy > -1; yy--) {
        int row = f->pc->y + yy;
        if (row > f->H - 1 || row < 0) continue;
        int full = 1, gb = 0;
        for (int xx = 0; xx < f->W; xx++) {
            uint8_t c = f->cell[row * f->W + xx];
            if (c == 8) gb = 1;
            if (c == 0) { full = 0; break; }
        }
        if (full) {
            remove_line(f, row);
            yy++;
            lines++;
            if (gb) (*garbage_lines)++;
        }
    }
    return lines;
}

/* ------------------------------------------------------------------ randomizer (SRC/randomizer.cpp) */
/* :55-62 (1000/7 is integer division -> 142) */
static void rnd_reset(player_t *p) { for (int i = 0; i < 7; i++) p->weight[i] = 1000 / 7; }
/* :10-32 */
static int rnd_piece(player_t *p) {
    int pick = 0;
    float r = mt_unit_float(&p->piece_gen) * 1000;
    for (int i = 0; i < 7; i++) {
        r -= p->weight[i];
        if (r < 0) { pick = i; break; }
    }
    float adjust = (p->weight[pick] / 4) * 3;
    p->weight[pick] -= adjust;
    adjust = (float)((double)adjust / 6.0);
    for (int i = 0; i < 7; i++)
        if (i != pick) p->weight[i] += adjust;
    return pick;
}
/* :38-45 (noStack is never true on this path) */
static int rnd_hole(player_t *p, int W) { return (int16_t)(mt_unit_float(&p->hole_gen) * W); }

/* ------------------------------------------------------------------ DropDelay (SRC/DropDelay.cpp) */
static int dd_check(player_t *p, int32_t t) { /* :3-21 */
    if (t - p->inc_drop_delay_time > p->inc_drop_delay) {
        if (p->drop_delay > 200) p->drop_delay -= 10;
        else if (p->drop_delay > 100) p->drop_delay -= 5;
        else if (p->drop_delay > 50) p->drop_delay -= 2;
        else if (p->drop_delay > 10) p->drop_delay -= 1;
        p->inc_drop_delay_time = t;
    }
    if (t - p->drop_delay_time > p->drop_delay) { p->drop_delay_time = t; return 1; }
    return 0;
}
static void dd_reset(player_t *p, int32_t t) { p->drop_delay_time = t; p->lockdown = 0; }   /* :23-26 */
static void dd_clear(player_t *p) {                                                           /* :28-35 */
    p->inc_drop_delay = 3000; p->inc_drop_delay_time = 0; p->drop_delay = 1000;
    p->drop_delay_time = 0; p->lockdown_time = 0; p->lockdown = 0;
}
static void dd_set(player_t *p, int32_t t) { if (!p->lockdown) p->lockdown_time = t + 400; p->lockdown = 1; } /* :37-41 */
static int dd_check_lockdown(const player_t *p, int32_t t) { return p->lockdown && t > p->lockdown_time; }  /* :43-48 */

/* ------------------------------------------------------------------ GarbageHandler (SRC/Garbage.cpp) */
static void gq_pop_front(player_t *p) { memmove(p->q, p->q + 1, (size_t)(p->qn - 1) * sizeof(garb_t)); p->qn--; }
static int gq_count(const player_t *p) { int t = 0; for (int i = 0; i < p->qn; i++) t += p->q[i].count; return (uint16_t)t; } /* :9-14 */
static void gq_clear(player_t *p) { p->qn = 0; p->min_remaining = 1000; }                         /* :16-20 */
static void gq_add(player_t *p, uint16_t amount, int32_t t) {                                    /* :22-24 */
    if (p->qn >= GARB_CAP) abort();
    p->q[p->qn].count = (int16_t)amount; p->q[p->qn].delay = t + 1000; p->qn++;
}
static uint16_t gq_block(player_t *p, uint16_t amount, int32_t t, int freeze) {                  /* :26-52 */
    if (p->qn == 0) return amount;
    int32_t delay = p->q[0].delay;
    while (amount && p->qn) {
        p->q[0].count--;
        amount--;
        if (p->q[0].count == 0) gq_pop_front(p);
    }
    if (p->qn) {
        if (delay > p->q[0].delay) p->q[0].delay = delay;
        if (freeze) {
            int32_t a = p->q[0].delay + 450, b = t + p->min_remaining + 450;
            p->q[0].delay = a < b ? a : b;
        }
    } else p->min_remaining = 1000;
    return amount;
}
static int gq_check(player_t *p, int32_t t) {                                                    /* :54-72 */
    if (p->qn == 0) return 0;
    if (t > p->q[0].delay) {
        int32_t delay = p->q[0].delay + p->add_delay;
        if (--p->q[0].count == 0) gq_pop_front(p);
        if (p->qn) {
            if (delay > p->q[0].delay) p->q[0].delay = delay;
            p->min_remaining = p->q[0].delay - t;
        } else p->min_remaining = 1000;
        return 1;
    }
    int32_t rem = p->q[0].delay - t;
    if (rem < p->min_remaining) p->min_remaining = rem;
    return 0;
}

/* ------------------------------------------------------------------ ComboCounter (SRC/Combo.cpp) */
static void cb_clear(player_t *p) { p->combo_start = 0; p->combo_time = 0; p->combo_count = 0; p->line_count = 0; } /* :7-13 */
static void cb_increase(player_t *p, int32_t t, uint8_t amount) {                                /* :15-30 */
    if (p->combo_count == 0) { p->combo_start = t; p->combo_time = 0; }
    p->combo_count++;
    float line_time = 0;
    for (int i = 0; i < amount; i++) {
        p->line_count++;
        line_time += 1000 / p->line_count;
    }
    p->combo_time = (int32_t)((float)p->combo_time + ((float)(800 / p->combo_count) + line_time));
}
static uint16_t cb_check(player_t *p, int32_t t) {                                               /* :32-48 */
    int32_t left = p->combo_start + p->combo_time - t;
    p->remaining = left < 0 ? 0 : (uint16_t)left;
    if (t > p->combo_start + p->combo_time && p->combo_count != 0) {
        float duration = 1.f + t / 60000.f * 0.1f;
        uint16_t sent = (uint16_t)(pow(p->combo_count, 1.4 + p->combo_count * 0.01) * duration);
        p->combo_count = 0;
        p->line_count = 0;
        return sent;
    }
    return 0;
}
static void cb_no_clear(player_t *p) { p->combo_time -= 200; }                                   /* :50-52 */

/* ------------------------------------------------------------------ GamePlay (SRC/gamePlay.cpp) */
static field_t field_of(const tor_vec *v, player_t *p) { field_t f = {p->cell, &p->cur, v->H, v->W}; return f; }

static void copy_piece(const tor_vec *v, player_t *p, uint8_t kind) {                            /* :84-88 */
    p->cur = v->base[kind];
    p->cur.x = (int8_t)((v->W - 4) / 2);
    p->cur.y = 0;
}
static int make_new_piece(const tor_vec *v, player_t *p) {                                       /* :71-82 */
    copy_piece(v, p, p->nextpiece);
    p->nextpiece = v->map[rnd_piece(p)];
    field_t f = field_of(v, p);
    if (!fits(&f)) { stamp(&f); return 1; }
    return 0;
}
static int send_lines(player_t *p, int lines) {                                                  /* :160-171 */
    p->data_lines_cleared = (uint16_t)(p->data_lines_cleared + lines);
    if (lines == 0) { cb_no_clear(p); return 0; }
    int amount = gq_block(p, (uint16_t)(lines - 1), p->time_ms, 1);
    cb_increase(p, p->time_ms, (uint8_t)lines);
    return amount;
}
static void hd_make(const tor_vec *v, player_t *p) {                                             /* :48-52 */
    field_t f = field_of(v, p);
    drop(&f); stamp(&f); dd_reset(p, p->time_ms);
}
static int hd_finish(const tor_vec *v, player_t *p) {                                            /* :54-59 */
    field_t f = field_of(v, p);
    int gl;
    int sent = send_lines(p, clear_lines(&f, &gl));
    if (make_new_piece(v, p)) return -1;
    return sent;
}
static int hd_full(const tor_vec *v, player_t *p) {                                              /* :38-46 */
    hd_make(v, p);
    return hd_finish(v, p);
}
static int gp_move_down(const tor_vec *v, player_t *p) {                                         /* :61-69 */
    field_t f = field_of(v, p);
    if (move_down(&f)) { dd_reset(p, p->time_ms); return 1; }
    dd_set(p, p->time_ms);
    return 0;
}
/* :199-204.  :202 writes columns 0..9 whatever W is; for W<10 that runs past the row (for the bottom
 * row: past the vector, into allocator slack) so the visible effect is "all W columns"; for W>10 the
 * columns >=10 keep what the (unshifted) bottom row held.  hole==W (u==1.0f) writes one byte past the
 * vector: no visible effect, the row simply has no hole. */
static void add_garbage_line(const tor_vec *v, player_t *p, uint8_t hole) {
    int H = v->H, W = v->W;
    memmove(p->cell, p->cell + W, (size_t)(H - 1) * (size_t)W);
    for (int xx = 0; xx < 10 && xx < W; xx++) p->cell[(H - 1) * W + xx] = 8;
    if (hole < W) p->cell[(H - 1) * W + hole] = 0;
}
static int push_garbage(const tor_vec *v, player_t *p) {                                         /* :179-192 */
    add_garbage_line(v, p, (uint8_t)rnd_hole(p, v->W));
    field_t f = field_of(v, p);
    if (p->cur.y > 0) p->cur.y--;
    if (!fits(&f)) {
        if (p->cur.y > 0) p->cur.y--;
        else return 1;
    }
    return 0;
}
static int delay_check(const tor_vec *v, player_t *p, int elapsed) {                             /* :90-114 */
    p->time_ms += elapsed;
    if (dd_check(p, p->time_ms)) gp_move_down(v, p);
    if (dd_check_lockdown(p, p->time_ms) && !gp_move_down(v, p)) return hd_full(v, p);
    int add = 0;
    while (p->incoming_lines >= 1) { add++; p->incoming_lines -= 1.f; }
    if (add) gq_add(p, (uint16_t)add, p->time_ms);
    uint16_t sent = cb_check(p, p->time_ms);
    if (sent) sent = gq_block(p, sent, p->time_ms, 0);
    if (gq_check(p, p->time_ms))
        if (push_garbage(v, p)) return -1;
    return sent;
}
static void restart_round(const tor_vec *v, player_t *p) {                                       /* :206-216 */
    memset(p->cell, 0, (size_t)v->H * (size_t)v->W);
    p->cur.kind = 7;
    gq_clear(p);
    cb_clear(p);
    p->data_lines_cleared = 0;
    dd_clear(p);
    p->time_ms = 0;
    p->incoming_lines = 0;
    p->lines_cleared = 0;
    p->dead = 0;
}
static void seed_player(const tor_vec *v, player_t *p, uint32_t time_value) {                    /* :218-230 */
    uint32_t s = (uint32_t)(int32_t)(int16_t)(uint16_t)time_value; /* randomizer.cpp:34-36,47-49: short */
    mt_seed(&p->hole_gen, s);
    mt_seed(&p->piece_gen, s);
    rnd_reset(p);
    make_new_piece(v, p);
    if (!v->only_zs)
        while (p->nextpiece == 2 || p->nextpiece == 3) { rnd_reset(p); make_new_piece(v, p); }
    make_new_piece(v, p);
}

/* ------------------------------------------------------------------ PythonHandle (SRC/PythonHandle.cpp) */
static void distribute(const tor_vec *v, game_t *g, int sender, int amount) {                    /* :124-136 */
    float others = (float)(v->np - 1);
    if (others < 1) return;
    float each = amount / others;
    for (int i = 0; i < v->np; i++)
        if (i != sender) g->pl[i].incoming_lines += each;
}
static void apply_key(const tor_vec *v, player_t *p, int key) {                                  /* :73-112 */
    field_t f = field_of(v, p);
    switch (key) {
    case 1: move_left(&f); break;
    case 2: while (move_left(&f)) {} break;
    case 3: move_right(&f); break;
    case 4: while (move_right(&f)) {} break;
    case 5: gp_move_down(v, p); break;
    case 6: while (gp_move_down(v, p)) {} break;
    case 7: hd_make(v, p); break;
    case 8: rot_cw(&f); break;
    case 9: rot_ccw(&f); break;
    case 10: rot_180(&f); break;
    default: break;
    }
}
/* :149-188.  action_finish (:114-122) returns true only on death: after distributeLines it falls off
 * the end of the function, which in both validated builds of the reference reads as false. */
static int finish_actions(const tor_vec *v, game_t *g, int elapsed) {
    if (g->round_over) return 1;
    for (int i = 0; i < v->np; i++) {
        player_t *p = &g->pl[i];
        if (p->dead) continue;
        int sent = hd_finish(v, p);
        if (sent == -1) { p->dead = 1; break; }
        if (sent) distribute(v, g, i, sent);
    }
    int alive = 0;
    for (int i = 0; i < v->np; i++) {
        player_t *p = &g->pl[i];
        if (p->dead) continue;
        int sent = delay_check(v, p, elapsed);
        if (sent == -1) { p->dead = 1; continue; }
        if (sent) distribute(v, g, i, sent);
        alive++;
        p->reward = (uint8_t)(p->data_lines_cleared - p->lines_cleared);
        p->lines_cleared = p->data_lines_cleared;
        p->incoming_lines_count = (uint8_t)gq_count(p);
    }
    if ((v->np > 1 && alive < 2) || !alive) { g->round_over = 1; return 1; }
    return 0;
}
static void reset_game(const tor_vec *v, game_t *g, uint32_t time_value) {                       /* :49-71 */
    g->round_over = 0;
    int winner = -1, alive = 0;
    for (int i = 0; i < v->np; i++) {
        if (!g->pl[i].dead) { alive++; winner = i; }
        restart_round(v, &g->pl[i]);
    }
    g->last_winner = winner;
    if (v->np == 1) g->last_winner = 0;
    if (alive > 1) g->last_winner = -1;
    for (int i = 0; i < v->np; i++) seed_player(v, &g->pl[i], time_value);
    g->env_done = 0; /* tetris_environment.reset, tetris_environment.py:66-68 */
}

/* ------------------------------------------------------------------ TestField (SRC/TestField.cpp) */
typedef struct {
    uint8_t tok[PATH_CAP];
    int n;
} path_t;

typedef struct {
    field_t f;
    piece_t pc;
    path_t path;        /* test_path */
    path_t best_path;   /* finesseMove.path */
    int best_x, best_rot;
    /* output */
    uint8_t *seq;       /* [cap][lcap] */
    int *seq_len;
    int cap, lcap, count, overflow;
} tf_t;

static void path_push(path_t *p, int t) { if (p->n < PATH_CAP) p->tok[p->n++] = (uint8_t)t; }

static int tf_key_of(int tok) { /* convert, :3-35 */
    switch (tok) {
    case 255: return 3;
    case 254: return 1;
    case 253: return 6;
    case 252: return 5;
    case 241: return 8;
    case 242: return 10;
    case 243: return 9;
    default: return tok;
    }
}
static void tf_emit_begin(tf_t *t) {
    if (t->count >= t->cap) { t->overflow = 1; return; }
    t->seq_len[t->count] = 0;
}
static void tf_emit_tok(tf_t *t, int tok) {
    if (t->count >= t->cap) return;
    int n = t->seq_len[t->count];
    if (n < t->lcap) t->seq[(size_t)t->count * (size_t)t->lcap + (size_t)n] = (uint8_t)tf_key_of(tok);
    else t->overflow = 1;
    t->seq_len[t->count] = n + 1;
}
static void tf_emit_end(tf_t *t) { if (t->count < t->cap) t->count++; }
static void tf_emit_rotation(tf_t *t, int delta) { /* addRotationValue, :37-42 */
    if (delta < 0) delta += 4;
    if (delta) tf_emit_tok(t, 240 + delta);
}
static void tf_emit_lateral(tf_t *t, int x) { /* makeStartSequence, :44-49 */
    int spawn = (t->f.W - 4) / 2;
    if (x > spawn) for (int i = 0; i < x - spawn; i++) tf_emit_tok(t, 255);
    else for (int i = 0; i < spawn - x; i++) tf_emit_tok(t, 254);
}

static uint8_t tf_move_up(tf_t *t) { /* :392-410 */
    int blocked_below = 0;
    t->pc.y++;
    if (!fits(&t->f)) blocked_below = 1;
    t->pc.y--;
    uint8_t count = 0;
    do { t->pc.y--; count++; } while (fits(&t->f));
    count--;
    t->pc.y++;
    if (count && blocked_below) return 253;
    return count;
}
static int tf_record(tf_t *t) { /* setFinesseMove, :166-172 */
    t->best_path = t->path;
    t->best_x = t->pc.x;
    t->best_rot = t->pc.current_rotation;
    return 1;
}
static int tf_try_left(tf_t *t, int clear) { /* :202-220 */
    for (;;) {
        if (!move_left(&t->f)) return 0;
        if (clear) t->path.n = 0;
        path_push(&t->path, 255);
        uint8_t up = tf_move_up(t);
        if (up) path_push(&t->path, up);
        if (t->pc.y == 0) return tf_record(t);
        clear = 0;
    }
}
static int tf_try_right(tf_t *t, int clear) { /* :222-238 */
    for (;;) {
        if (!move_right(&t->f)) return 0;
        if (clear) t->path.n = 0;
        path_push(&t->path, 254);
        uint8_t up = tf_move_up(t);
        if (up) path_push(&t->path, up);
        if (t->pc.y == 0) return tf_record(t);
        clear = 0;
    }
}
static int tf_try_up(tf_t *t, int turn) { /* :240-259 */
    t->path.n = 0;
    uint8_t up = tf_move_up(t);
    path_push(&t->path, turn + 240);
    if (up) path_push(&t->path, up);
    if (t->pc.y == 0) return tf_record(t);
    if (tf_try_left(t, 0)) return 1;
    t->path.n = 0;
    path_push(&t->path, turn + 240);
    if (up) path_push(&t->path, up);
    return tf_try_right(t, 0);
}
static int tf_wall_kick(tf_t *t) { /* doWallKick, :280-356 */
    const piece_t start = t->pc;
    int r, working = 0;
    for (r = 0; r < 4; r++) {
        working = 1;
        if (r == start.current_rotation) { working = 0; continue; }
        while (t->pc.current_rotation != r) piece_rcw(&t->pc);
        if (fits(&t->f)) {
            int turn = (start.current_rotation - r) & 3;
            int ok = tf_try_up(t, turn);
            t->pc = start;
            return ok;
        }
        piece_t *p = &t->pc;
        p->y--; if (fits(&t->f)) break;
        p->y++; p->x--; if (fits(&t->f)) break;
        p->x += 2; if (fits(&t->f)) break;
        p->y--; p->x -= 2; if (fits(&t->f)) break;
        p->x += 2; if (fits(&t->f)) break;
        p->y++; p->x -= 3; if (fits(&t->f)) break;
        p->x += 4; if (fits(&t->f)) break;
        p->x -= 2;
        working = 0;
    }
    if (!working) { t->pc = start; return 0; }
    int turn = (start.current_rotation - r) & 3;
    const piece_t pre = t->pc;
    if (turn == 1) rot_cw(&t->f);
    else if (turn == 2) rot_180(&t->f);
    else if (turn == 3) rot_ccw(&t->f);
    else { t->pc = start; return 0; }
    if (t->pc.x != start.x || t->pc.y != start.y) { t->pc = start; return 0; }
    t->pc = pre;
    if (tf_try_up(t, turn)) { t->pc = start; return 1; }
    t->pc = pre;
    t->path.n = 0;
    path_push(&t->path, 240 + turn);
    if (tf_try_left(t, 0)) { t->pc = start; return 1; }
    t->pc = pre;
    if (t->path.n > 1) { t->path.n = 0; path_push(&t->path, 240 + turn); }
    int ok = tf_try_right(t, 0);
    t->pc = start;
    return ok;
}
static void tf_flip_in_place(tf_t *t) { /* r180KeepPos, :358-390 */
    piece_t *p = &t->pc;
    piece_rcw(p); piece_rcw(p);
    int s = (p->kind == 2) ? -1 : 1; /* S mirrors the I/Z adjustment */
    if (p->kind == 4 || p->kind == 3 || p->kind == 2) {
        switch (p->current_rotation) {
        case 0: p->x = (int8_t)(p->x + s); break;
        case 1: p->y = (int8_t)(p->y + s); break;
        case 2: p->x = (int8_t)(p->x - s); break;
        default: p->y = (int8_t)(p->y - s); break;
        }
    }
}
static int tf_reverse_kick(tf_t *t) { /* reverseWallkick, :261-278 */
    if (t->pc.kind == 6) return 0;
    if (move_right(&t->f)) { move_left(&t->f); return 0; }
    if (move_left(&t->f)) { move_right(&t->f); return 0; }
    const piece_t start = t->pc;
    if (tf_wall_kick(t)) { t->pc = start; return 1; }
    if (t->pc.kind == 2 || t->pc.kind == 3 || t->pc.kind == 4) {
        tf_flip_in_place(t);
        int ok = tf_wall_kick(t);
        t->pc = start;
        return ok;
    }
    t->pc = start;
    return 0;
}
static int tf_reachable(tf_t *t) { /* finesseIsPossible, :189-200 — note: no restore between the failed
                                      left search and the right search */
    const piece_t start = t->pc;
    if (tf_reverse_kick(t)) { t->pc = start; return 1; }
    t->pc = start;
    if (tf_try_left(t, 1)) { t->pc = start; return 1; }
    int ok = tf_try_right(t, 1);
    t->pc = start;
    return ok;
}
static void tf_emit_finesse(tf_t *t) { /* useFinesseMove (use_mask==2 branch), :127-164 */
    tf_emit_begin(t);
    tf_emit_rotation(t, t->best_rot - t->pc.rotation);
    tf_emit_lateral(t, t->best_x);
    for (int i = t->best_path.n - 1; i >= 0; i--) {
        int tok = t->best_path.tok[i];
        if (tok < 240) for (int k = 0; k < tok; k++) tf_emit_tok(t, 252);
        else tf_emit_tok(t, tok);
    }
    tf_emit_tok(t, 7);
    tf_emit_end(t);
}
static void tf_cavities(tf_t *t) { /* tryAllFinesseMoves, :174-187 */
    const piece_t rest = t->pc;
    for (int yy = rest.y + 2; yy < t->f.H - 1; yy++) {
        t->pc = rest;
        t->pc.y = (int8_t)yy;
        if (fits(&t->f)) {
            drop(&t->f);
            yy = t->pc.y;
            if (tf_reachable(t)) tf_emit_finesse(t);
        }
    }
    t->pc = rest;
}
static void tf_column(tf_t *t) { /* findNextMove, :113-125 */
    tf_emit_begin(t);
    tf_emit_rotation(t, t->pc.current_rotation - t->pc.rotation);
    tf_emit_lateral(t, t->pc.x);
    tf_emit_tok(t, 7);
    tf_emit_end(t);
    drop(&t->f);
    tf_cavities(t);
}
/* GamePlay::getMask(2) (gamePlay.cpp:232-239) + TestField::setPiece/getMask (:58-111) */
static void tf_enumerate(const tor_vec *v, const player_t *p, tf_t *t) {
    t->f.cell = p->cell; t->f.pc = &t->pc; t->f.H = v->H; t->f.W = v->W;
    t->pc = p->cur;
    t->pc.x = (int8_t)((v->W - 4) / 2);
    t->pc.y = 0;
    t->path.n = 0; t->best_path.n = 0; t->count = 0; t->overflow = 0;
    int kind = t->pc.kind;
    int nrot = (kind == 6) ? 1 : (kind == 4 || kind == 2 || kind == 3) ? 2 : 4;
    for (int xx = -1; xx < v->W - 1; xx++)
        for (int r = 0; r < nrot; r++) {
            t->pc.x = (int8_t)xx;
            t->pc.y = 0;
            if (kind != 6)
                while (t->pc.current_rotation != r) piece_rcw(&t->pc);
            if (!fits(&t->f)) continue;
            tf_column(t);
        }
}

/* ------------------------------------------------------------------ exported API */
tor_vec *tor_create(int n_games, int n_players, int height, int width, const uint8_t piece_map[7]) {
    if (n_players < 1 || n_players > TOR_MAX_PLAYERS || n_games < 1) return NULL;
    tor_vec *v = calloc(1, sizeof *v);
    v->n = n_games; v->np = n_players; v->H = height; v->W = width;
    memcpy(v->map, piece_map, 7);
    v->only_zs = 1; /* set_pieces, PythonHandle.h:116-121 */
    for (int i = 0; i < 7; i++) if (piece_map[i] != 2 && piece_map[i] != 3) v->only_zs = 0;
    make_base_pieces(v->base);
    v->g = calloc((size_t)n_games, sizeof(game_t));
    v->cells = calloc((size_t)n_games * (size_t)n_players * (size_t)height * (size_t)width + 16, 1);
    for (int i = 0; i < n_games; i++)
        for (int p = 0; p < n_players; p++) {
            player_t *pl = &v->g[i].pl[p];
            pl->cell = v->cells + ((size_t)i * (size_t)n_players + (size_t)p) * (size_t)height * (size_t)width;
            pl->add_delay = 450;  /* GarbageHandler ctor, Garbage.cpp:7 */
            pl->remaining = 0;    /* Combo.h:27 is uninitialised in the reference; harness ignores it until first finish */
            rnd_reset(pl);
            restart_round(v, pl); /* PythonHandle::init, PythonHandle.cpp:15-18 */
        }
    return v;
}
void tor_destroy(tor_vec *v) { if (!v) return; free(v->cells); free(v->g); free(v); }
int tor_n_games(const tor_vec *v) { return v->n; }

void tor_reset(tor_vec *v, const int32_t *idx, int n, const uint32_t *seeds) {
    for (int i = 0; i < n; i++) reset_game(v, &v->g[idx ? idx[i] : i], seeds[i]);
}

void tor_step(tor_vec *v, const int32_t *idx, int n, const uint8_t *player, const uint8_t *keys,
              const uint8_t *key_len, int lmax, int time_ms, int finalize, int8_t *reward_base,
              uint8_t *done, uint8_t *lines) {
    for (int i = 0; i < n; i++) {
        game_t *g = &v->g[idx ? idx[i] : i];
        int pl = player[i];
        /* make_actions, PythonHandle.cpp:138-147: every other player receives the null action [0] */
        if (!g->round_over && !g->pl[pl].dead)
            for (int k = 0; k < key_len[i]; k++) apply_key(v, &g->pl[pl], keys[(size_t)i * (size_t)lmax + (size_t)k]);
        if (finalize) g->env_done = (uint8_t)finish_actions(v, g, time_ms);
        /* reward_fcn, tetris_environment.py:135-149 (extra_rewards=False) */
        int base = 0;
        if (g->env_done) {
            /* backend.states[1 - player] is a PYTHON index (tetris_environment.py:139): for players beyond the second it counts
             * from the end of the list — player 2 of 3 reads itself, player 3 of 4 reads player 2 */
            int opp = 1 - pl >= 0 ? 1 - pl : v->np + (1 - pl);
            int me = g->pl[pl].dead != 0, you = (v->np > 1) ? (g->pl[opp].dead != 0) : 0;
            base = you - me;
            if (me && you) base = -1;
        }
        if (reward_base) reward_base[i] = (int8_t)base;
        if (done) done[i] = g->env_done;
        if (lines) for (int p = 0; p < v->np; p++) lines[(size_t)i * (size_t)v->np + (size_t)p] = g->pl[p].reward;
    }
}

static int piece_index_of(const piece_t *pc) { /* state_processors.py:24,34,43: col_code[piece.max()] */
    static const int code[8] = {-1, 5, 4, 1, 0, 2, 6, 3};
    int mx = 0;
    for (int i = 0; i < 16; i++) if (((const uint8_t *)pc->grid)[i] > mx) mx = ((const uint8_t *)pc->grid)[i];
    return code[mx & 7];
}

void tor_observe(const tor_vec *v, const int32_t *idx, int n, uint8_t *field, uint8_t *vec, uint8_t *piece_idx) {
    size_t hw = (size_t)v->H * (size_t)v->W;
    for (int i = 0; i < n; i++) {
        const game_t *g = &v->g[idx ? idx[i] : i];
        for (int p = 0; p < v->np; p++) {
            const player_t *pl = &g->pl[p];
            size_t o = (size_t)i * (size_t)v->np + (size_t)p;
            if (field) for (size_t c = 0; c < hw; c++) field[o * hw + c] = pl->cell[c] > 0;
            int kind = piece_index_of(&pl->cur);
            if (vec) {
                uint8_t *q = vec + o * TOR_VEC_LEN;
                q[0] = (uint8_t)pl->cur.x;
                q[1] = (uint8_t)pl->cur.y;
                q[2] = pl->incoming_lines_count;
                uint16_t ct = (uint16_t)(pl->remaining + 50); /* uint16 numpy arithmetic wraps */
                q[3] = (uint8_t)((ct < 25000 ? ct : 25000) / 100);
                q[4] = pl->combo_count;
                for (int k = 0; k < 7; k++) q[5 + k] = (pl->nextpiece == k);
                for (int k = 0; k < 7; k++) q[12 + k] = (kind == k);
            }
            if (piece_idx) piece_idx[o] = (uint8_t)kind;
        }
    }
}

void tor_raw(const tor_vec *v, const int32_t *idx, int n, uint8_t *field, uint8_t *grid, int32_t *scal) {
    size_t hw = (size_t)v->H * (size_t)v->W;
    for (int i = 0; i < n; i++) {
        const game_t *g = &v->g[idx ? idx[i] : i];
        for (int p = 0; p < v->np; p++) {
            const player_t *pl = &g->pl[p];
            size_t o = (size_t)i * (size_t)v->np + (size_t)p;
            if (field) memcpy(field + o * hw, pl->cell, hw);
            if (grid) memcpy(grid + o * 16, pl->cur.grid, 16);
            if (scal) {
                int32_t *s = scal + o * 8;
                s[0] = pl->cur.x; s[1] = pl->cur.y; s[2] = pl->incoming_lines_count; s[3] = pl->remaining;
                s[4] = pl->combo_count; s[5] = pl->nextpiece; s[6] = pl->reward; s[7] = pl->dead;
            }
        }
    }
}

void tor_hidden(const tor_vec *v, const int32_t *idx, int n, int32_t *out) {
    for (int i = 0; i < n; i++) {
        const game_t *g = &v->g[idx ? idx[i] : i];
        for (int p = 0; p < v->np; p++) {
            const player_t *pl = &g->pl[p];
            int32_t *s = out + ((size_t)i * (size_t)v->np + (size_t)p) * 16;
            s[0] = pl->time_ms; s[1] = pl->drop_delay; s[2] = pl->drop_delay_time; s[3] = pl->inc_drop_delay_time;
            s[4] = pl->lockdown_time; s[5] = pl->lockdown; s[6] = pl->qn; s[7] = gq_count(pl);
            s[8] = pl->min_remaining; s[9] = (int32_t)(pl->incoming_lines * 1000.f); s[10] = pl->lines_cleared;
            s[11] = pl->qn ? pl->q[0].count : 0; s[12] = pl->qn ? pl->q[0].delay : 0;
            s[13] = pl->combo_start; s[14] = pl->combo_time; s[15] = pl->line_count;
        }
    }
}

int tor_enumerate(tor_vec *v, const int32_t *idx, int n, const uint8_t *player, int bar_null_moves,
                  int kmax, int lmax, int32_t *K, uint8_t *keys, uint8_t *key_len, int32_t *K_raw) {
    enum { RAW_CAP = 1024, RAW_L = 256 };
    tf_t *t = calloc(1, sizeof *t);
    uint8_t *raw = malloc((size_t)RAW_CAP * RAW_L);
    int *raw_len = malloc(sizeof(int) * RAW_CAP);
    int *keep = malloc(sizeof(int) * (RAW_CAP + 1));
    t->seq = raw; t->seq_len = raw_len; t->cap = RAW_CAP; t->lcap = RAW_L;
    int bad = 0;
    for (int i = 0; i < n; i++) {
        const game_t *g = &v->g[idx ? idx[i] : i];
        tf_enumerate(v, &g->pl[player[i]], t);
        if (K_raw) K_raw[i] = t->count;
        /* action_list.__init__, data_types/action_list.py:4-20.  keep[] holds raw indices, -1 = null action [0].
         * The raw list never contains [0] (every sequence ends in 7), so the null action is always prepended. */
        int nk = 0;
        keep[nk++] = -1;
        for (int a = 0; a < t->count; a++) {
            int dup = 0;
            for (int b = 1; b < nk && !dup; b++) {
                int c = keep[b];
                dup = raw_len[c] == raw_len[a] && !memcmp(raw + (size_t)c * RAW_L, raw + (size_t)a * RAW_L, (size_t)raw_len[a]);
            }
            if (!dup) keep[nk++] = a;
        }
        int first = (bar_null_moves && nk > 1) ? 1 : 0; /* remove_nulls: only while len > 1 */
        int kk = nk - first, trunc = t->overflow;
        if (kk > kmax) { kk = kmax; trunc = 1; }
        K[i] = nk - first;
        for (int a = 0; a < kk; a++) {
            int c = keep[first + a];
            uint8_t *dst = keys + ((size_t)i * (size_t)kmax + (size_t)a) * (size_t)lmax;
            memset(dst, 0, (size_t)lmax);
            if (c < 0) { key_len[(size_t)i * (size_t)kmax + (size_t)a] = 1; continue; }
            int len = raw_len[c];
            if (len > lmax) { len = lmax; trunc = 1; }
            memcpy(dst, raw + (size_t)c * RAW_L, (size_t)len);
            key_len[(size_t)i * (size_t)kmax + (size_t)a] = (uint8_t)len;
        }
        for (int a = kk; a < kmax; a++) {
            memset(keys + ((size_t)i * (size_t)kmax + (size_t)a) * (size_t)lmax, 0, (size_t)lmax);
            key_len[(size_t)i * (size_t)kmax + (size_t)a] = 0;
        }
        bad += trunc;
    }
    free(keep); free(raw_len); free(raw); free(t);
    return bad;
}

void tor_copy_game(tor_vec *to, int dst, const tor_vec *from, int src) {
    size_t hw = (size_t)to->H * (size_t)to->W;
    game_t *d = &to->g[dst];
    const game_t *s = &from->g[src];
    for (int p = 0; p < to->np; p++) {
        uint8_t *cell = d->pl[p].cell;
        memcpy(cell, s->pl[p].cell, hw);
        d->pl[p] = s->pl[p];
        d->pl[p].cell = cell;
    }
    d->round_over = s->round_over;
    d->last_winner = s->last_winner;
    d->env_done = s->env_done;
}

int tor_expand(tor_vec *v, const int32_t *idx, int n, const uint8_t *player, int bar_null_moves, int finalize,
               int time_ms, int kmax, int lmax, int32_t *K, uint8_t *keys, uint8_t *key_len, uint8_t *field,
               uint8_t *vec, uint8_t *piece_idx, int8_t *reward_base, uint8_t *done, uint8_t *lines) {
    tor_vec *s = tor_create(1, v->np, v->H, v->W, v->map); /* the "anchor"-restored working copy */
    size_t hw = (size_t)v->H * (size_t)v->W, P = (size_t)v->np;
    int bad = 0;
    for (int i = 0; i < n; i++) {
        int32_t gi = idx ? idx[i] : i;
        uint8_t *kb = keys + (size_t)i * (size_t)kmax * (size_t)lmax, *lb = key_len + (size_t)i * (size_t)kmax;
        bad += tor_enumerate(v, &gi, 1, player + i, bar_null_moves, kmax, lmax, K + i, kb, lb, NULL);
        int kk = K[i] < kmax ? K[i] : kmax;
        for (int a = 0; a < kk; a++) {
            size_t o = (size_t)i * (size_t)kmax + (size_t)a;
            tor_copy_game(s, 0, v, gi);
            tor_step(s, NULL, 1, player + i, kb + (size_t)a * (size_t)lmax, lb + a, lmax, time_ms, finalize,
                     reward_base ? reward_base + o : NULL, done ? done + o : NULL, lines ? lines + o * P : NULL);
            tor_observe(s, NULL, 1, field ? field + o * P * hw : NULL, vec ? vec + o * P * TOR_VEC_LEN : NULL,
                        piece_idx ? piece_idx + o * P : NULL);
        }
    }
    tor_destroy(s);
    return bad;
}

void tor_poke_dead(tor_vec *v, int game, int player, int dead) { v->g[game].pl[player].dead = (uint8_t)dead; }
void tor_poke_cell(tor_vec *v, int game, int player, int y, int x, int value) {
    v->g[game].pl[player].cell[y * v->W + x] = (uint8_t)value;
}

void tor_piece_stream(uint32_t seed, const uint8_t piece_map[7], int n, uint8_t *out) {
    tor_vec *v = tor_create(1, 1, 22, 10, piece_map);
    player_t *p = &v->g[0].pl[0];
    seed_player(v, p, seed);
    if (n > 0) out[0] = p->cur.kind;
    for (int i = 1; i < n; i++) { out[i] = p->nextpiece; p->nextpiece = v->map[rnd_piece(p)]; }
    tor_destroy(v);
}
void tor_hole_stream(uint32_t seed, int width, int n, uint8_t *out) {
    static const uint8_t ident[7] = {0, 1, 2, 3, 4, 5, 6};
    tor_vec *v = tor_create(1, 1, 22, width, ident);
    player_t *p = &v->g[0].pl[0];
    seed_player(v, p, seed);
    for (int i = 0; i < n; i++) out[i] = (uint8_t)rnd_hole(p, width);
    tor_destroy(v);
}
void tor_piece_grids(uint8_t *out) {
    piece_t base[7];
    make_base_pieces(base);
    for (int k = 0; k < 7; k++) {
        piece_t p = base[k];
        for (int i = 0; i < 4; i++) {
            memcpy(out + ((size_t)k * 4 + p.current_rotation) * 16, p.grid, 16);
            piece_rcw(&p);
        }
    }
}
