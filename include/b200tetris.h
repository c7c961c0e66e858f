/* b200tetris — C ABI of the B200-native batched Tetris engine.
 *
 * Drop-in boundary for DRL-Tetris' `environment/game_backend` (pybind11 module `tetris_env`) as driven by
 * `environment/tetris_environment_vector.py`.  One handle = a vector of independent 2-player games resident in the
 * HBM of one GPU; every call launches hand-written sm_100a kernels on the caller's stream.  Citations below are
 * file:line under the reference tree (/root/reference).
 *
 * Conventions
 *   - every function returns B2T_OK (0) or a negative b2t_status; b2t_last_error() gives the text; nothing throws.
 *   - pointers named d_* are DEVICE pointers on the handle's device, 16-byte aligned; pointers named h_* are HOST
 *     pointers.  The library never takes ownership of caller memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).  Device-pointer calls are asynchronous;
 *     *_host calls run on the legacy default stream (plus two internal streams for large batches) and synchronise before
 *     returning; a caller that also drives the handle from non-blocking streams of its own orders the two itself.
 *   - game_idx == NULL means "games 0..n-1"; otherwise n distinct game ids (duplicates in one call are undefined,
 *     like stepping one PythonHandle from two threads).  Ids must lie in [0, n_games): a *_host call answers B2T_ERR_ARG for a
 *     list that holds any other id (the reference raises IndexError); a device-pointer call cannot look at the list, so its
 *     kernels treat such a row as inert — no game is read or written, its outputs are zeros / untouched, its K is 0.
 *   - h_* arrays may have any alignment; the zero-copy path of b2t_step_host is only taken when every array is naturally
 *     aligned for its element type and pinned over its whole length.
 *   - a handle is not thread-safe (neither is the reference); handles on different devices are independent.
 *   - array shapes below are written for two players ([rows, 2, ...]); a handle created with another n_players has that many
 *     entries on the player axis.
 *   - there is NO CPU fallback: without a CUDA device b2t_create fails with B2T_ERR_CUDA.
 */
#ifndef B200TETRIS_H
#define B200TETRIS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2T_ABI_VERSION 3
#define B2T_VEC_LEN 19  /* x,y,incoming_lines,combo_time,combo_count,nextpiece[7],piece[7]  (state_processors.py:33-41) */
#define B2T_SNAPSHOT_SLOTS 4

typedef enum {
    B2T_OK = 0,
    B2T_ERR_ARG = -1,          /* bad argument (NULL, range, alignment) */
    B2T_ERR_UNSUPPORTED = -2,  /* field size not instantiated (csrc/b2t_geoms.h), n_players outside 1..4, or a two-player-only feature */
    B2T_ERR_CUDA = -3,         /* CUDA runtime error (no device, launch failure, out of memory) */
    B2T_ERR_CAPACITY = -4,     /* caller buffer too small for the result */
    B2T_ERR_STATE = -5         /* call order (e.g. step by placement index without a prior enumerate) */
} b2t_status;

/* Settings of tetris_environment (experiments/presets.py:123-182 "default") that reach the backend. */
typedef struct {
    int32_t n_games;             /* n_envs of tetris_environment_vector (tetris_environment_vector.py:10-17) */
    int32_t n_players;           /* settings["n_players"]: 1..4 (PythonHandle.cpp:5-25).  2 is what every preset of the reference uses and
                                    what the fast kernels serve; 1, 3 and 4 run on a plain compatibility kernel with the same rules:
                                    axis 1 of every per-player array then has n_players entries, observations are B2T_U8 in
                                    player-index order, keys are B2T_KEYS_U8, and b2t_get_game / b2t_set_game / B2T_FLAG_TILES
                                    (whose structures describe two players) answer B2T_ERR_UNSUPPORTED */
    int32_t height, width;       /* settings["game_size"] = [H, W] (PythonHandle.cpp:5-8) */
    uint8_t piece_map[7];        /* (settings["pieces"]*7)[:7]  (tetris_environment.py:191-193 -> set_pieces) */
    uint8_t bar_null_moves;      /* settings["bar_null_moves"] (action_list remove_null) */
    int32_t time_ms_per_action;  /* settings["time_elapsed_each_action"], 400 */
    int32_t piece_stream_len;    /* draws tabulated per seed (0 = default 4096); a game drawing more raises a flag */
    int32_t hole_stream_len;     /* garbage holes tabulated per seed (0 = default 512) */
    int32_t device;              /* CUDA device ordinal */
    uint32_t flags;              /* B2T_FLAG_* */
    uint32_t seed;               /* base of the engine-chosen round seeds (b2t_step_args.auto_reset) */
} b2t_config;

/* Keep the raw cell values of PythonHandle.states[p].field (1..7 = piece.tile, 8 = garbage; gameField.cpp:105-110,
 * gamePlay.cpp:202) as three extra bit planes per player, for the renderer (environment/env_utils/draw_tetris.py:80-101) and
 * raw-field parity: b2t_get_game / b2t_set_game then exchange 0..8 instead of 0/1.  Off by default: nothing observable
 * through state_dict needs it, it adds 6*ceil(H*W/32) words of HBM traffic per stepped game and steps run the generic kernel. */
#define B2T_FLAG_TILES 1u

typedef struct b2t_handle b2t_handle;

/* geometry-derived constants of a handle */
typedef struct {
    int32_t kmax;           /* capacity of a placement list */
    int32_t lmax;           /* capacity of a key sequence */
    int32_t state_bytes;    /* bytes of packed state per game in HBM */
    int32_t obs_bytes;      /* n_players*(H*W + 19 + 1) */
    int32_t tile_bytes;     /* bytes of colour planes per game (0 without B2T_FLAG_TILES) */
} b2t_info;

/* tetris_env.set_pieces + tetris_env.PythonHandle(n_players,[H,W]) x n_games (tetris_environment.py:34-39).
 * Games start un-reset (no piece); call b2t_reset before stepping, as tetris_environment.__init__ does (:41). */
int b2t_create(const b2t_config *cfg, b2t_handle **out);
int b2t_destroy(b2t_handle *h);
const char *b2t_last_error(const b2t_handle *h); /* h may be NULL: error of the last failed b2t_create */
int b2t_get_info(const b2t_handle *h, b2t_info *out);

/* PythonHandle.reset() (PythonHandle.cpp:49-71; tetris_environment.py:66-75; vector :55-58).
 * d_seeds[i] plays the role of time(NULL): only its low 16 bits matter (randomizer.cpp:34-36,47-49). */
int b2t_reset(b2t_handle *h, const int32_t *d_game_idx, int32_t n, const uint32_t *d_seeds, void *stream);

/* Observation tensors written by step / observe / expand: the reference's state_dict for BOTH players
 * (state_processors.py:23-54; layout consumed by agents/agent_utils/state_unpack.py:88-105).  Any pointer may be NULL. */
typedef enum { B2T_U8 = 0, B2T_F16 = 1, B2T_BF16 = 2, B2T_F32 = 3, B2T_BITS = 4 } b2t_dtype;

typedef struct {
    void *d_field;         /* [rows, 2, H, W]  field>0 */
    void *d_vec;           /* [rows, 2, 19]    x, y, incoming_lines, min(25000,combo_time+50)//100, combo_count, nextpiece 1-hot, piece 1-hot */
    uint8_t *d_piece_idx;  /* [rows, 2] */
    int32_t *d_scalars;    /* [rows, 2, 8] raw State views (PythonHandle.h:54-82): x,y,inc_lines,combo_time,combo_count,nextpiece,reward,dead */
    /* Player order of axis 1 of d_field / d_vec / d_piece_idx.  0: player index (0, 1).  1: perspective order (me, opponent),
     * the layout agents/agent_utils/state_unpack.py:35-58,88-137 feeds the networks ([..., :12] of d_vec is its `vector`,
     * d_field its `visual`, d_piece_idx its `piece`): "me" is the acting player for b2t_step / b2t_expand and d_me[row] for
     * b2t_observe.  d_scalars always stays in player-index order. */
    int32_t perspective;
    const uint8_t *d_me;   /* [rows], only read by b2t_observe when perspective == 1 */
    /* Element type of d_field and d_vec (b2t_dtype).  B2T_U8 is the reference's state_dict dtype; B2T_F16 / B2T_BF16 / B2T_F32
     * write the same values (exactly representable: 0/1 cells, integers <= 255) as the floating type the consumer's network
     * takes, so no conversion pass re-reads the tensors (state_unpack.py:88-105 casts them on the host in the reference).
     * B2T_BITS is an extra, compact form for consumers that unpack on the device: d_field is uint32 [rows, 2, SW], SW = ceil(H*W/32),
     * bit (r*W + c) % 32 of word (r*W + c) / 32 = cell (r, c) of that player's field>0 (unused high bits of the last word are 0);
     * d_vec stays uint8.  One eighth of the bytes of the uint8 cells. */
    int32_t dtype;
} b2t_obs;

/* tetris_environment.perform_action for n games at once (tetris_environment.py:102-116; vector :79-86):
 * PythonHandle.make_action (the acting player's keys, null action for the other; PythonHandle.cpp:138-147,73-112)
 * then, if finalize, PythonHandle.finish_action(time_ms) (PythonHandle.cpp:149-188). */
typedef struct {
    const int32_t *d_game_idx;     /* NULL or [n] */
    int32_t n;
    const uint8_t *d_player;       /* [n] acting player 0/1 */
    /* action, one of: */
    const uint8_t *d_keys;         /* [n, lmax] key codes 0..10 (others ignored, PythonHandle.cpp:109-110) */
    const uint8_t *d_key_len;      /* [n], or NULL: rows are zero-terminated (key 0 is a no-op, PythonHandle.cpp:109-110, so the
                                      length is the position of the last non-zero key) */
    int32_t lmax;
    const int32_t *d_placement;    /* [n] index into the list of the last b2t_enumerate/b2t_expand of that game */
    int32_t finalize;              /* simulate/finalize switch of perform_action; 1 = normal step */
    int32_t time_ms;               /* <=0: cfg.time_ms_per_action */
    const uint32_t *d_reset_seeds; /* optional [n]: games that end in this call are reset with these seeds right
                                      after their reward/done are produced (fused worker-loop reset, worker.py:157-160) */
    /* results */
    int8_t *d_reward;              /* [n] base term of reward_fcn: 0, or youdead-medead when done (tetris_environment.py:135-144) */
    uint8_t *d_done;               /* [n] */
    uint8_t *d_lines;              /* [n,2] states[p].reward = lines cleared since the previous finish (PythonHandle.cpp:176) */
    b2t_obs obs;                   /* state after the call (after the fused reset, if any) */
    int32_t key_format;            /* B2T_KEYS_U8 (one key code per byte) or B2T_KEYS_PACKED4 (4-bit tokens, two per byte, low nibble
                                      first; a row is lmax/2 bytes, lmax a multiple of 8; codes 11..15 are no-ops like any other) */
    int32_t auto_reset;            /* d_reset_seeds == NULL and auto_reset != 0: games that end are reset with an engine-chosen
                                      seed — a counter hash of (cfg.seed, game id, rounds started), the role time(NULL) plays in
                                      PythonHandle::seed — so no seeds travel to the device */
} b2t_step_args;
#define B2T_KEYS_U8 0
#define B2T_KEYS_PACKED4 1
int b2t_step(b2t_handle *h, const b2t_step_args *a, void *stream);

/* get_state + state_dict (tetris_environment.py:131-133; vector :99-101) */
int b2t_observe(b2t_handle *h, const int32_t *d_game_idx, int32_t n, const b2t_obs *obs, void *stream);

/* get_actions: TestField::getMask(2) + action_list de-duplication / null handling
 * (tetris_environment.py:77-85, TestField.cpp:64-187, data_types/action_list.py:4-20).
 * The lists stay cached per game (for d_placement / b2t_expand).  d_K [n]; optional d_keys [n,kmax,lmax], d_key_len [n,kmax]. */
int b2t_enumerate(b2t_handle *h, const int32_t *d_game_idx, int32_t n, const uint8_t *d_player, int32_t *d_K,
                  uint8_t *d_keys, uint8_t *d_key_len, int32_t kmax, int32_t lmax, void *stream);

/* simulate_all_actions (tetris_environment.py:87-100,127-129): enumerate, then every placement replayed on a scratch
 * copy of its game (finalize as in perform_action) — the all-placements world model.  Compact output: placement j of
 * board i sits at row d_offsets[i]+j; rows >= capacity are dropped and reported through *d_total > capacity. */
typedef struct {
    const int32_t *d_game_idx;
    int32_t n;
    const uint8_t *d_player;   /* [n] */
    int32_t finalize;
    int32_t time_ms;
    int32_t capacity;          /* rows available in the output arrays */
    int32_t *d_K;              /* [n] */
    int32_t *d_offsets;        /* [n+1] exclusive prefix sum of K; d_offsets[n] = total */
    int32_t *d_board;          /* optional [capacity] board i of each row */
    uint8_t *d_keys;           /* optional [capacity, lmax] */
    uint8_t *d_key_len;        /* optional [capacity] */
    int32_t lmax;
    int8_t *d_reward;          /* optional [capacity] */
    uint8_t *d_done;           /* optional [capacity] */
    uint8_t *d_lines;          /* optional [capacity,2] */
    b2t_obs obs;               /* rows = placements */
} b2t_expand_args;
int b2t_expand(b2t_handle *h, const b2t_expand_args *a, void *stream);

/* The select step of the actor loop (sventon_agent: evaluate every placement, act on the best): for every board of the last
 * b2t_expand, the placement with the highest score, first one on ties like numpy.argmax.  d_scores [total] float32 in the row
 * order of b2t_expand; d_placement [n] is what b2t_step takes as d_placement.  Asynchronous, no host round trip: enumerate ->
 * expand -> (caller's network) -> select -> step can be captured in one CUDA graph.  `capacity` is the capacity that b2t_expand
 * was given (= rows of d_scores): rows it dropped (offsets beyond capacity) have no score and are never chosen; a board left
 * without any row takes placement 0. */
int b2t_select(b2t_handle *h, const float *d_scores, const int32_t *d_offsets, int32_t n, int32_t capacity, int32_t *d_placement,
               void *stream);

/* PythonHandle.copy() / PythonHandle.set() for the whole vector (PythonHandle.cpp:36-42; tetris_environment.py:165-176) */
int b2t_snapshot(b2t_handle *h, int32_t slot, void *stream);
int b2t_restore(b2t_handle *h, int32_t slot, void *stream);

/* Pickling equivalent (PythonHandle.h:283-308; drl_tetris/runner.py:69-88): packed state of game g, state_bytes each. */
int b2t_export_state(b2t_handle *h, int32_t game, void *h_buf, void *stream);
int b2t_import_state(b2t_handle *h, int32_t game, const void *h_buf, void *stream);

/* The same for n games in one call (get_state / set of a whole vector of envs): h_buf is [n][state_bytes]; h_game_idx NULL =
 * games 0..n-1.  Synchronise. */
int b2t_export_states(b2t_handle *h, const int32_t *h_game_idx, int32_t n, void *h_buf);
int b2t_import_states(b2t_handle *h, const int32_t *h_game_idx, int32_t n, const void *h_buf);

/* ---- one game in the reference's own terms: what PythonHandle.states[p] shows (PythonHandle.h:54-82) plus the hidden
 * members its pickle support walks (PythonHandle.h:207-308).  This is the exchange format of the per-env compatibility
 * facade (b200tetris/tetris_env.py: a PythonHandle-shaped object per game slot), of pickling, and of white-box tests. */
typedef struct {
    /* State views */
    uint8_t piece[16];          /* field.piece.grid: 4x4, colour code of the piece (1..7, state_processors.py:24) or 0 */
    int32_t x, y;               /* field.piece.posX/posY */
    int32_t inc_lines;          /* incoming_lines_count */
    int32_t combo_time;         /* combo.remaining (uint16) */
    int32_t combo_count;        /* combo.comboCount */
    int32_t nextpiece;
    int32_t reward;             /* lines cleared by the last finish_action */
    int32_t dead;
    /* hidden */
    int32_t piece_kind;         /* field.piece.piece 0..6; 7 = no piece yet (never-reset game) */
    int32_t rotation;           /* field.piece.current_rotation */
    int32_t time_ms;            /* GamePlay::time_ms */
    int32_t drop_delay, drop_delay_time, increase_drop_delay_time, lockdown_time, lockdown;   /* DropDelay.h:6-18 */
    int32_t combo_start, combo_time_total, line_count;                                        /* Combo.h:8-33 comboStart, comboTime, lineCount */
    int32_t min_remaining;      /* GarbageHandler::minRemaining */
    int32_t incoming_lines;     /* GamePlay::incoming_lines (float in the reference; integral with two players) */
    int32_t lines_pending;      /* data.linesCleared - linesCleared */
    int32_t garbage_n;          /* GarbageHandler::garbage.size() (at most 6 are kept) */
    int32_t garbage_count[6];
    int32_t garbage_delay[6];
    int32_t piece_draws, hole_draws;   /* position in the player's piece / hole stream (rander) */
} b2t_player_raw;

typedef struct {
    int32_t seed16;             /* the (short) seed of the current round (randomizer.cpp:34-36,47-49) */
    int32_t round_over;         /* PythonHandle::round_over */
    int32_t done;               /* tetris_environment.done as maintained by b2t_step */
    int32_t last_winner;        /* PythonHandle::last_winner */
    int32_t overflow;           /* bit 0: garbage queue outgrew 6 entries, bit 1: a stream table was outrun */
    int32_t rounds_started;     /* resets so far (mod 65536): the counter behind engine-chosen round seeds (auto_reset) */
    b2t_player_raw player[2];
} b2t_game_raw;

/* h_field: optional [2, H, W] bytes, the players' fields (non-zero = occupied; the raw values 1..8 with B2T_FLAG_TILES).  b2t_set_game with h_field == NULL keeps the
 * boards.  `piece` is output only (the piece is set through piece_kind / rotation).  Both synchronise. */
int b2t_get_game(b2t_handle *h, int32_t game, b2t_game_raw *out, uint8_t *h_field);
int b2t_set_game(b2t_handle *h, int32_t game, const b2t_game_raw *in, const uint8_t *h_field);
/* PythonHandle.copy() / set() for single games: game `src` of handle `from` into game `dst` of handle `to` (same field size,
 * same device; `from` may be `to`).  Asynchronous on `stream`. */
int b2t_copy_game(b2t_handle *to, int32_t dst, b2t_handle *from, int32_t src, void *stream);

/* The writable numpy views the front-end pokes (data_types/state.py:9-17 dead[0]=1; scripts/eval.py:156 field[:,:]=0) */
int b2t_poke_dead(b2t_handle *h, int32_t game, int32_t player, int32_t dead, void *stream);
int b2t_clear_field(b2t_handle *h, int32_t game, int32_t player, void *stream);

/* Episode counters accumulated by b2t_step since creation, for an end-of-rollout reduction across ranks:
 * [0] env-steps [1] episodes ended [2] wins player0 [3] wins player1 [4] both dead [5] lines cleared
 * [6] games whose garbage queue overflowed its 6 entries [7] games that outran a stream table.  h_out[8]. */
int b2t_stats(b2t_handle *h, uint64_t *h_out, void *stream);

/* ---- host-buffer entry points (what a ctypes / pybind11 shim that owns numpy arrays calls) ----
 * Same semantics as above with HOST pointers.  Pageable buffers are copied through pinned staging owned by the handle, pinned
 * ones directly; b2t_step_host of a large batch (>= 65 536 rows) whose buffers are ALL pinned and that asks for no observation
 * tensors copies nothing: the kernel reads and writes the caller's pinned pages (zero-copy). */
int b2t_reset_host(b2t_handle *h, const int32_t *h_game_idx, int32_t n, const uint32_t *h_seeds);
int b2t_step_host(b2t_handle *h, const int32_t *h_game_idx, int32_t n, const uint8_t *h_player, const uint8_t *h_keys,
                  const uint8_t *h_key_len, int32_t lmax, const int32_t *h_placement, int32_t finalize, int32_t time_ms,
                  const uint32_t *h_reset_seeds, int8_t *h_reward, uint8_t *h_done, uint8_t *h_lines,
                  uint8_t *h_field, uint8_t *h_vec, uint8_t *h_piece_idx, int32_t *h_scalars, int32_t key_format, int32_t auto_reset);
int b2t_observe_host(b2t_handle *h, const int32_t *h_game_idx, int32_t n, uint8_t *h_field, uint8_t *h_vec,
                     uint8_t *h_piece_idx, int32_t *h_scalars);
int b2t_enumerate_host(b2t_handle *h, const int32_t *h_game_idx, int32_t n, const uint8_t *h_player, int32_t *h_K,
                       uint8_t *h_keys, uint8_t *h_key_len, int32_t kmax, int32_t lmax);
/* rows written = min(total, capacity); *h_total = total */
int b2t_expand_host(b2t_handle *h, const int32_t *h_game_idx, int32_t n, const uint8_t *h_player, int32_t finalize,
                    int32_t time_ms, int32_t capacity, int32_t *h_K, int32_t *h_offsets, int32_t *h_total,
                    uint8_t *h_keys, uint8_t *h_key_len, int32_t lmax, int8_t *h_reward, uint8_t *h_done, uint8_t *h_lines,
                    uint8_t *h_field, uint8_t *h_vec, uint8_t *h_piece_idx, int32_t *h_scalars /* optional [capacity,2,8] */);

/* Pinned host memory for the *_host entry points (buffers allocated here, or pinned by the caller in any other way, are copied
 * without the staging pass).  write_combined != 0 asks for write-combined pages: slow for the host to READ, but the device's
 * reads of them skip the CPU cache snoop — use it for inputs the host only writes (keys, players). */
int b2t_host_alloc(size_t bytes, int32_t write_combined, void **out);
int b2t_host_free(void *p);

/* Measurement aid (bench.py roofline.mix_peak): one launch with exactly the HBM traffic of a fused b2t_step of games 0..n-1 — state
 * vectors read and written back with one word changed (the handle's games are garbage afterwards), key rows / players / lengths / seeds read, reward / done / lines / observation tensors
 * written with arbitrary bytes — and no game logic: the bandwidth this read/write mix reaches on the device.  All pointers are
 * device pointers shaped as in b2t_step_args (d_keys [n,lmax]); the caller's output arrays are overwritten with garbage.
 * step_like != 0: with the step kernel's shared-memory footprint (its 16 resident warps per SM); 0: as many warps as fit. */
int b2t_debug_traffic(b2t_handle *h, int32_t n, const uint8_t *d_keys, int32_t lmax, const uint8_t *d_player,
                      const uint8_t *d_key_len, const uint32_t *d_seeds, int8_t *d_reward, uint8_t *d_done, uint8_t *d_lines,
                      uint8_t *d_field, uint8_t *d_vec, uint8_t *d_piece_idx, int32_t step_like, void *stream);

/* Stream-table probe for tests: first n entries of the piece / hole stream of a 16-bit seed. */
int b2t_debug_streams(b2t_handle *h, uint32_t seed16, int32_t n, uint8_t *h_pieces, uint8_t *h_holes);

#ifdef __cplusplus
}
#endif
#endif /* B200TETRIS_H */
