/* TEST INFRASTRUCTURE ONLY — never linked, imported or executed by the product path.
 *
 * Plain-C CPU restatement of the DRL-Tetris game backend (environment/game_backend/source/)
 * and of the thin Python semantics layered on it by environment/tetris_environment.py,
 * environment/data_types/action_list.py and environment/env_utils/state_processors.py.
 *
 * PARITY PIN: this restatement is checked against the compiled, unmodified reference
 * (oracle/_ref/{A,B}/tetris_env*.so, built by oracle/Makefile) by tests/test_oracle_vs_reference.py
 * and against the committed fixtures in tests/golden/ (generated from that compiled reference by
 * tests/golden/make_golden.py).  The reference itself ships no tests or golden vectors.
 *
 * It deliberately keeps the reference's formulation (byte-per-cell field, 4x4 piece grids,
 * per-player std::mt19937 streams, std::vector-like garbage queue) so that it is an independent
 * check of the bit-board / table-driven CUDA engine.
 */
#ifndef TETRIS_ORACLE_H
#define TETRIS_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TOR_MAX_PLAYERS 4
#define TOR_VEC_LEN 19 /* x,y,incoming_lines,combo_time,combo_count,nextpiece[7],piece[7] */

typedef struct tor_vec tor_vec; /* a vector of independent games */

tor_vec *tor_create(int n_games, int n_players, int height, int width, const uint8_t piece_map[7]);
void tor_destroy(tor_vec *v);
int tor_n_games(const tor_vec *v);

/* PythonHandle::reset (PythonHandle.cpp:49-66) with time(NULL) == seeds[i]. idx==NULL -> games 0..n-1 */
void tor_reset(tor_vec *v, const int32_t *idx, int n, const uint32_t *seeds);

/* tetris_environment.perform_action (tetris_environment.py:102-116): make_action with the null action
 * for every other player, then (finalize) finish_action(time_ms).  keys: [n, lmax] u8, key_len: [n].
 * reward_base: youdead-medead rule of reward_fcn (:135-149); done: env.done after the call.
 * lines: [n, n_players] u8 = states[p].reward after the call. */
void tor_step(tor_vec *v, const int32_t *idx, int n, const uint8_t *player, const uint8_t *keys,
              const uint8_t *key_len, int lmax, int time_ms, int finalize, int8_t *reward_base,
              uint8_t *done, uint8_t *lines);

/* state_dict (state_processors.py:23-54) for every player: field [n,P,H,W], vec [n,P,19], piece_idx [n,P] */
void tor_observe(const tor_vec *v, const int32_t *idx, int n, uint8_t *field, uint8_t *vec,
                 uint8_t *piece_idx);

/* Raw PythonHandle.states[p] views (PythonHandle.h:54-82): field with tile colours [n,P,H,W],
 * piece grid [n,P,16], and scalars [n,P,8] as int32: x,y,inc_lines,combo_time,combo_count,nextpiece,reward,dead */
void tor_raw(const tor_vec *v, const int32_t *idx, int n, uint8_t *field, uint8_t *grid, int32_t *scal);

/* Hidden state the reference exposes only through __getstate__ (PythonHandle.h:207-281), [n,P,16] int32:
 * 0 time_ms 1 dropDelay 2 dropDelayTime 3 increaseDropDelayTime 4 lockdownTime 5 lockdown
 * 6 garbage entries 7 garbage total 8 minRemaining 9 incoming_lines*1000 10 linesCleared
 * 11 front.count 12 front.delay 13 comboStart 14 comboTime 15 lineCount */
void tor_hidden(const tor_vec *v, const int32_t *idx, int n, int32_t *out);

/* get_actions (tetris_environment.py:77-85): TestField::getMask(2) raw list, then action_list()
 * de-duplication and null handling (action_list.py:4-20).  K [n]; keys [n,kmax,lmax]; key_len [n,kmax];
 * K_raw [n] (length of masks[p].action before action_list) may be NULL.  Returns the number of games
 * whose K exceeded kmax or whose longest sequence exceeded lmax (those entries are truncated). */
int tor_enumerate(tor_vec *v, const int32_t *idx, int n, const uint8_t *player, int bar_null_moves,
                  int kmax, int lmax, int32_t *K, uint8_t *keys, uint8_t *key_len, int32_t *K_raw);

/* simulate_all_actions (tetris_environment.py:87-100,127-129): get_actions, then for every action restore
 * the anchor, perform_action(simulate=True, finalize), snapshot.  Padded outputs: field [n,kmax,P,H,W],
 * vec [n,kmax,P,19], piece_idx [n,kmax,P], reward_base/done [n,kmax], lines [n,kmax,P]; any may be NULL.
 * The games themselves are left untouched.  Return value as tor_enumerate. */
int tor_expand(tor_vec *v, const int32_t *idx, int n, const uint8_t *player, int bar_null_moves, int finalize,
               int time_ms, int kmax, int lmax, int32_t *K, uint8_t *keys, uint8_t *key_len, uint8_t *field,
               uint8_t *vec, uint8_t *piece_idx, int8_t *reward_base, uint8_t *done, uint8_t *lines);

/* PythonHandle.copy()/set(): deep copy game src of `from` into game dst of `to` (same geometry) */
void tor_copy_game(tor_vec *to, int dst, const tor_vec *from, int src);

/* state.lock()-style write access used by the front-end (data_types/state.py:9-17) and tests */
void tor_poke_dead(tor_vec *v, int game, int player, int dead);
void tor_poke_cell(tor_vec *v, int game, int player, int y, int x, int value);

/* RNG restatement exposed for table tests: first `n` values of the piece stream (mapped, after the
 * GamePlay::seed procedure, gamePlay.cpp:218-230: out[0]=first current piece, out[1]=first nextpiece, ...)
 * and of the hole stream (randomizer.cpp:38-45) for time()==seed. */
void tor_piece_stream(uint32_t seed, const uint8_t piece_map[7], int n, uint8_t *out);
void tor_hole_stream(uint32_t seed, int width, int n, uint8_t *out);

/* Spawn-orientation grids per piece and rotation, [7][4][16] (pieces.cpp:5-53, gamePlay.cpp:116-158) */
void tor_piece_grids(uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
