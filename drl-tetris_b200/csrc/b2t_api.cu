// b200tetris — C ABI (include/b200tetris.h): handle lifetime, argument checking, launch plumbing, pinned staging for
// the *_host entry points.  No game logic lives here and there is no CPU path: every call ends in a kernel launch.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200tetris.h"
#include "b2t_geoms.h"
#define B2T_COMMON_KERNELS
#include "b2t_kernels.cuh"

using namespace b2t;

#define X(h, w) const GeomOps* b2t_geom_ops_##h##_##w();
B2T_FOR_EACH_GEOMETRY(X)
#undef X

static const GeomOps* find_geom(int H, int W) {
#define X(h, w) if (H == h && W == w) return b2t_geom_ops_##h##_##w();
    B2T_FOR_EACH_GEOMETRY(X)
#undef X
    return nullptr;
}

static thread_local std::string g_create_error;

struct Arena {  // grow-only pinned host + device scratch for the *_host calls
    unsigned char* h = nullptr;
    unsigned char* d = nullptr;
    size_t cap = 0, at = 0;
    cudaError_t reserve(size_t bytes) {
        at = 0;
        if (bytes <= cap) return cudaSuccess;
        if (h) cudaFreeHost(h);
        if (d) cudaFree(d);
        h = d = nullptr; cap = 0;
        bytes = (bytes + (1u << 20)) & ~((size_t)(1u << 20) - 1);
        cudaError_t e = cudaMallocHost((void**)&h, bytes);
        if (e != cudaSuccess) return e;
        e = cudaMalloc((void**)&d, bytes);
        if (e != cudaSuccess) return e;
        cap = bytes;
        return cudaSuccess;
    }
    size_t take(size_t bytes) { size_t o = at; at += (bytes + 255) & ~(size_t)255; return o; }
    void release() { if (h) cudaFreeHost(h); if (d) cudaFree(d); h = d = nullptr; cap = 0; }
};

struct b2t_handle {
    b2t_config cfg;
    const GeomOps* ops = nullptr;
    int np = 2, nv = 0;             // players per game; uint4 vectors per game (ops->NV with two players)
    bool multi() const { return np != 2; }   // served by gamen_kernel (any player count) instead of the two-player kernels
    uint4* state = nullptr;
    size_t state_vecs = 0;
    uint4* snap[B2T_SNAPSHOT_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t* plc = nullptr;
    int32_t* kcache = nullptr;
    int32_t* scan_sums = nullptr;
    int32_t* d_total = nullptr;
    int4* row_meta = nullptr;       // expand: {board slot, game, placement, player} of every output row
    size_t row_meta_cap = 0;
    uint8_t* piece_tab = nullptr;
    uint8_t* hole_tab = nullptr;
    double* powtab = nullptr;
    unsigned long long* stats = nullptr;
    uint4* xfer = nullptr;   // one game's vectors for export/import
    uint32_t* tiles = nullptr;          // colour planes (B2T_FLAG_TILES)
    uint32_t* snap_tiles[B2T_SNAPSHOT_SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t* xfer_tiles = nullptr;
    size_t tile_words_total = 0;
    DevTabs tabs;
    Arena in, out;
    cudaStream_t pipe[2] = {nullptr, nullptr};   // b2t_step_host slices large batches over two streams (copies overlap the kernel)
    bool enumerated = false;
    std::string err;
};

static int fail(b2t_handle* h, int code, const char* fmt, const char* a = "", long long b = 0) {
    char buf[512];
    snprintf(buf, sizeof buf, fmt, a, b);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}
#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) return fail(h, B2T_ERR_CUDA, "CUDA error: %s (line %lld)", cudaGetErrorString(e_), __LINE__); \
    } while (0)

constexpr int PIPE_MIN_ROWS = 1 << 16;   // b2t_step_host: batches at least this large are sliced over two streams
constexpr int PIPE_SLICES = 2;       // measured on the B200 box: H2D (25-38 GB/s pinned) bounds the call; 2 slices 0.68 ms, 1: 0.74, 8: 0.92 (profiles/e2e_probe.py)
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// every id of a caller's HOST game list must name a game of the handle (the kernels ignore rows whose id does not, but a host call
// answers with an error the way the reference's list indexing raises IndexError)
static bool ids_in_range(const int32_t* h_idx, int32_t n, int32_t n_games) {
    if (!h_idx) return true;
    uint32_t worst = 0;
    for (int32_t i = 0; i < n; i++) { const uint32_t v = (uint32_t)h_idx[i]; worst = v > worst ? v : worst; }
    return n <= 0 || worst < (uint32_t)n_games;
}
#define CK_IDS(name)                                                                                                   \
    do {                                                                                                               \
        if (!ids_in_range(h_idx, n, h->cfg.n_games)) return fail(h, B2T_ERR_ARG, name ": game id outside [0, n_games)%s"); \
    } while (0)
// Streams of a pipelined host call are drained on every way out, so no asynchronous copy still touches the caller's buffers or the
// arenas after the call has returned (an error return included)
struct StreamDrain {
    cudaStream_t a, b;
    ~StreamDrain() { if (a) cudaStreamSynchronize(a); if (b) cudaStreamSynchronize(b); }
};

extern "C" {

const char* b2t_last_error(const b2t_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int b2t_destroy(b2t_handle* h) {
    if (!h) return B2T_OK;
    cudaSetDevice(h->cfg.device);
    cudaFree(h->state);
    for (auto s : h->snap) cudaFree(s);
    for (auto s : h->snap_tiles) cudaFree(s);
    cudaFree(h->tiles); cudaFree(h->xfer_tiles);
    cudaFree(h->plc); cudaFree(h->kcache); cudaFree(h->scan_sums); cudaFree(h->d_total); cudaFree(h->row_meta);
    cudaFree(h->piece_tab); cudaFree(h->hole_tab); cudaFree(h->powtab); cudaFree(h->stats); cudaFree(h->xfer);
    h->in.release(); h->out.release();
    for (auto st : h->pipe) if (st) cudaStreamDestroy(st);
    delete h;
    return B2T_OK;
}

int b2t_create(const b2t_config* cfg, b2t_handle** out) {
    b2t_handle* h = nullptr;
    if (!cfg || !out) return fail(h, B2T_ERR_ARG, "b2t_create: NULL argument%s");
    *out = nullptr;
    if (cfg->n_games < 1) return fail(h, B2T_ERR_ARG, "b2t_create: n_games must be >= 1%s");
    if (cfg->n_players < 1 || cfg->n_players > MAX_PLAYERS) return fail(h, B2T_ERR_UNSUPPORTED, "b2t_create: n_players must be 1..4%s");
    if (cfg->n_players != 2 && (cfg->flags & B2T_FLAG_TILES)) return fail(h, B2T_ERR_UNSUPPORTED, "b2t_create: colour planes are built for two players only%s");
    for (int i = 0; i < 7; i++)
        if (cfg->piece_map[i] > 6) return fail(h, B2T_ERR_ARG, "b2t_create: piece_map entries must be 0..6%s");
    const GeomOps* ops = find_geom(cfg->height, cfg->width);
    if (!ops) return fail(h, B2T_ERR_UNSUPPORTED, "b2t_create: field size %s%lld not instantiated (csrc/b2t_geoms.h)", "H*1000+W=", cfg->height * 1000LL + cfg->width);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail(h, B2T_ERR_CUDA, "b2t_create: no CUDA device — this library has no CPU path%s");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(h, B2T_ERR_ARG, "b2t_create: bad device ordinal%s");
    h = new b2t_handle();
    h->cfg = *cfg;
    h->ops = ops;
    h->np = cfg->n_players;
    h->nv = h->np == 2 ? ops->NV : (h->np * ops->PW + 3) / 4;
    if (h->cfg.time_ms_per_action <= 0) h->cfg.time_ms_per_action = 400;
    if (h->cfg.piece_stream_len <= 0) h->cfg.piece_stream_len = 4096;
    if (h->cfg.hole_stream_len <= 0) h->cfg.hole_stream_len = 512;
    if (h->cfg.piece_stream_len < 4) h->cfg.piece_stream_len = 4;
#define CKC(call)                                                                        \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            fail(nullptr, B2T_ERR_CUDA, "b2t_create: CUDA error: %s (line %lld)", cudaGetErrorString(e_), __LINE__); \
            b2t_destroy(h);                                                              \
            return B2T_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)
    CKC(cudaSetDevice(cfg->device));
    CKC(ops->prepare());
    const size_t n = (size_t)cfg->n_games;
    h->state_vecs = (size_t)h->nv * 32 * ((n + 31) / 32);   // tiles of 32 games
    CKC(cudaMalloc((void**)&h->state, h->state_vecs * sizeof(uint4)));
    CKC(cudaMalloc((void**)&h->kcache, n * sizeof(int32_t)));
    CKC(cudaMemset(h->kcache, 0, n * sizeof(int32_t)));
    CKC(cudaMalloc((void**)&h->scan_sums, (size_t)SCAN_TILE * sizeof(int32_t)));
    CKC(cudaMalloc((void**)&h->d_total, 64));
    CKC(cudaMalloc((void**)&h->stats, STAT_SLICES * STAT_WORDS * sizeof(unsigned long long)));
    CKC(cudaMemset(h->stats, 0, STAT_SLICES * STAT_WORDS * sizeof(unsigned long long)));
    CKC(cudaMalloc((void**)&h->xfer, (size_t)h->nv * sizeof(uint4)));
    if (cfg->flags & B2T_FLAG_TILES) {
        h->tile_words_total = (size_t)ops->tile_words * 32 * ((n + 31) / 32);
        CKC(cudaMalloc((void**)&h->tiles, h->tile_words_total * sizeof(uint32_t)));
        CKC(cudaMemset(h->tiles, 0, h->tile_words_total * sizeof(uint32_t)));
        CKC(cudaMalloc((void**)&h->xfer_tiles, (size_t)ops->tile_words * sizeof(uint32_t)));
    }
    const size_t plen = (size_t)h->cfg.piece_stream_len, hlen = (size_t)h->cfg.hole_stream_len;
    CKC(cudaMalloc((void**)&h->piece_tab, plen * RNG_SEEDS));
    CKC(cudaMalloc((void**)&h->hole_tab, hlen * RNG_SEEDS));
    CKC(cudaMalloc((void**)&h->powtab, 256 * sizeof(double)));
    {
        double pw[256];
        for (int c = 0; c < 256; c++) pw[c] = pow((double)c, 1.4 + c * 0.01);  // Combo.cpp:41, host libm
        CKC(cudaMemcpy(h->powtab, pw, sizeof pw, cudaMemcpyHostToDevice));
    }
    {
        PieceTables pt;
        make_piece_tables(pt);
        memcpy(h->tabs.pcn, pt.pcn, sizeof pt.pcn);
        memcpy(h->tabs.pcc, pt.pcc, sizeof pt.pcc);
        h->tabs.spawn_rots = 0;
        for (int k = 0; k < 7; k++) h->tabs.spawn_rots |= (uint32_t)pt.spawn_rot[k] << (2 * k);
        h->tabs.piece = h->piece_tab; h->tabs.hole = h->hole_tab; h->tabs.powtab = h->powtab;
        h->tabs.piece_len = (uint32_t)plen; h->tabs.hole_len = (uint32_t)hlen;
    }
    {
        PieceMap pm;
        bool only_zs = true;  // set_pieces, PythonHandle.h:116-121
        for (int i = 0; i < 7; i++) { pm.m[i] = cfg->piece_map[i]; if (pm.m[i] != 2 && pm.m[i] != 3) only_zs = false; }
        pm.m[7] = 0;
        fill_tables_kernel<<<RNG_SEEDS / 64, 64>>>(h->piece_tab, (uint32_t)plen, h->hole_tab, (uint32_t)hlen, pm, only_zs ? 1 : 0, cfg->width);
        CKC(cudaGetLastError());
    }
    if (!h->multi()) CKC(ops->launch_init(h->state, cfg->n_games, 0));
    else {
        CKC(cudaMemset(h->state, 0, h->state_vecs * sizeof(uint4)));
        KParams p;
        memset(&p, 0, sizeof p);
        p.state = h->state; p.n_games = cfg->n_games; p.n = cfg->n_games; p.np = h->np; p.nv = h->nv; p.tabs = h->tabs;
        p.flags = KF_INIT | KF_STORE;
        CKC(ops->launch_gamen(p, cfg->n_games, 0));
    }
    CKC(cudaDeviceSynchronize());
    *out = h;
    return B2T_OK;
}

int b2t_get_info(const b2t_handle* h, b2t_info* out) {
    if (!h || !out) return B2T_ERR_ARG;
    out->kmax = h->ops->KMAX; out->lmax = h->ops->LMAX;
    out->state_bytes = h->nv * 16;
    out->obs_bytes = h->np * (h->ops->H * h->ops->W + B2T_VEC_LEN + 1);
    out->tile_bytes = h->tiles ? h->ops->tile_words * 4 : 0;
    return B2T_OK;
}

static void base_params(b2t_handle* h, KParams& p) {
    memset(&p, 0, sizeof p);
    p.state = h->state; p.tiles = h->tiles; p.n_games = h->cfg.n_games; p.tabs = h->tabs; p.stats = h->stats;
    p.plc = h->plc; p.kcache = h->kcache; p.elapsed = h->cfg.time_ms_per_action;
    p.np = h->np; p.nv = h->nv;
}
// the kernels of this handle's player count
static cudaError_t launch_game(b2t_handle* h, const KParams& p, int rows, cudaStream_t s) {
    return h->multi() ? h->ops->launch_gamen(p, rows, s) : h->ops->launch_game(p, rows, s);
}
static int check_obs(b2t_handle* h, const b2t_obs& o) {
    if (!aligned16(o.d_field) || !aligned16(o.d_vec) || !aligned16(o.d_piece_idx) || !aligned16(o.d_scalars))
        return fail(h, B2T_ERR_ARG, "observation pointers must be 16-byte aligned%s");
    if (o.dtype < B2T_U8 || o.dtype > B2T_BITS) return fail(h, B2T_ERR_ARG, "b2t_obs.dtype must be B2T_U8, B2T_F16, B2T_BF16, B2T_F32 or B2T_BITS%s");
    if (h->multi() && (o.dtype != B2T_U8 || o.perspective))
        return fail(h, B2T_ERR_UNSUPPORTED, "n_players != 2: observations are uint8 in player-index order%s");
    return B2T_OK;
}
static void set_obs(KParams& p, const b2t_obs& o) {
    p.field = (uint8_t*)o.d_field; p.vec = (uint8_t*)o.d_vec; p.piece_idx = o.d_piece_idx; p.scalars = o.d_scalars;
    p.perspective = o.perspective ? 1 : 0; p.me = o.d_me; p.obs_dtype = o.dtype;
}

int b2t_reset(b2t_handle* h, const int32_t* d_idx, int32_t n, const uint32_t* d_seeds, void* stream) {
    if (!h) return B2T_ERR_ARG;
    if (n < 0 || n > h->cfg.n_games || !d_seeds) return fail(h, B2T_ERR_ARG, "b2t_reset: bad n or NULL seeds%s");
    CK(cudaSetDevice(h->cfg.device));
    KParams p; base_params(h, p);
    p.idx = d_idx; p.n = n; p.seeds = d_seeds; p.flags = KF_RESET | KF_STORE;
    CK(launch_game(h, p, n, (cudaStream_t)stream));
    return B2T_OK;
}

static int step_launch(b2t_handle* h, const b2t_step_args* a, int32_t game0, void* stream);
int b2t_step(b2t_handle* h, const b2t_step_args* a, void* stream) { return step_launch(h, a, 0, stream); }
static int step_launch(b2t_handle* h, const b2t_step_args* a, int32_t game0, void* stream) {
    if (!h || !a) return B2T_ERR_ARG;
    if (a->n < 0 || game0 < 0 || a->n + game0 > h->cfg.n_games || !a->d_player) return fail(h, B2T_ERR_ARG, "b2t_step: bad n or NULL d_player%s");
    if (!a->d_placement && (!a->d_keys || a->lmax < 1)) return fail(h, B2T_ERR_ARG, "b2t_step: need d_placement or d_keys+lmax%s");
    if (a->key_format != B2T_KEYS_U8 && a->key_format != B2T_KEYS_PACKED4) return fail(h, B2T_ERR_ARG, "b2t_step: bad key_format%s");
    if (a->key_format == B2T_KEYS_PACKED4 && (a->lmax & 7)) return fail(h, B2T_ERR_ARG, "b2t_step: packed keys need lmax to be a multiple of 8%s");
    if (h->multi() && a->key_format != B2T_KEYS_U8) return fail(h, B2T_ERR_UNSUPPORTED, "n_players != 2: keys are one code per byte%s");
    if (a->d_placement && !h->enumerated) return fail(h, B2T_ERR_STATE, "b2t_step: d_placement needs a prior b2t_enumerate/b2t_expand%s");
    if (int r = check_obs(h, a->obs)) return r;
    CK(cudaSetDevice(h->cfg.device));
    KParams p; base_params(h, p);
    p.idx = a->d_game_idx; p.game0 = game0; p.n = a->n; p.player = a->d_player;
    p.keys = a->d_keys; p.key_len = a->d_key_len; p.lmax = a->lmax; p.placement = a->d_placement;
    p.key_fmt = a->key_format; p.auto_reset = a->auto_reset ? 1 : 0; p.seed_base = h->cfg.seed;
    p.flags = KF_STEP | KF_STORE | (a->finalize ? KF_FINALIZE : 0u);
    if (a->time_ms > 0) p.elapsed = a->time_ms;
    p.reset_seeds = a->d_reset_seeds;
    p.reward = a->d_reward; p.done = a->d_done; p.lines = a->d_lines;
    set_obs(p, a->obs);
    CK(launch_game(h, p, a->n, (cudaStream_t)stream));
    return B2T_OK;
}

int b2t_observe(b2t_handle* h, const int32_t* d_idx, int32_t n, const b2t_obs* obs, void* stream) {
    if (!h || !obs) return B2T_ERR_ARG;
    if (n < 0 || n > h->cfg.n_games) return fail(h, B2T_ERR_ARG, "b2t_observe: bad n%s");
    if (int r = check_obs(h, *obs)) return r;
    if (obs->perspective && !obs->d_me) return fail(h, B2T_ERR_ARG, "b2t_observe: perspective order needs d_me%s");
    CK(cudaSetDevice(h->cfg.device));
    KParams p; base_params(h, p);
    p.idx = d_idx; p.n = n; p.flags = 0;
    set_obs(p, *obs);
    CK(launch_game(h, p, n, (cudaStream_t)stream));
    return B2T_OK;
}

static int ensure_cache(b2t_handle* h) {
    if (h->plc) return B2T_OK;
    CK(cudaMalloc((void**)&h->plc, (size_t)h->cfg.n_games * h->ops->KMAX * h->ops->KEYW * sizeof(uint32_t)));
    return B2T_OK;
}

int b2t_enumerate(b2t_handle* h, const int32_t* d_idx, int32_t n, const uint8_t* d_player, int32_t* d_K, uint8_t* d_keys,
                  uint8_t* d_key_len, int32_t kmax, int32_t lmax, void* stream) {
    if (!h) return B2T_ERR_ARG;
    if (n < 0 || n > h->cfg.n_games || !d_player) return fail(h, B2T_ERR_ARG, "b2t_enumerate: bad n or NULL d_player%s");
    if ((d_keys || d_key_len) && (kmax < 1 || lmax < 1)) return fail(h, B2T_ERR_ARG, "b2t_enumerate: kmax/lmax must be positive%s");
    CK(cudaSetDevice(h->cfg.device));
    if (int r = ensure_cache(h)) return r;
    EParams e;
    memset(&e, 0, sizeof e);
    e.state = h->state; e.n_games = h->cfg.n_games; e.idx = d_idx; e.n = n; e.player = d_player;
    e.nv = h->nv; e.np = h->np;
    e.bar_null = h->cfg.bar_null_moves ? 1 : 0; e.plc = h->plc; e.kcache = h->kcache; e.K = d_K; e.stats = h->stats; e.tabs = h->tabs;
    CK(h->ops->launch_enum(e, (cudaStream_t)stream));
    h->enumerated = true;
    if (d_keys || d_key_len) CK(h->ops->launch_unpack(h->plc, h->kcache, d_idx, h->cfg.n_games, n, kmax, lmax, d_keys, d_key_len, (cudaStream_t)stream));
    return B2T_OK;
}

static int scan_offsets(b2t_handle* h, const int32_t* d_K, int n, int32_t* d_offsets, cudaStream_t s) {
    const int tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (tiles > SCAN_TILE) return fail(h, B2T_ERR_ARG, "b2t_expand: n too large for the offset scan%s");
    scan_tile_sums<<<tiles, SCAN_TPB, 0, s>>>(d_K, n, h->scan_sums);
    scan_sums<<<1, SCAN_TPB, 0, s>>>(h->scan_sums, tiles);
    scan_finish<<<tiles, SCAN_TPB, 0, s>>>(d_K, n, h->scan_sums, d_offsets);
    CK(cudaGetLastError());
    return B2T_OK;
}

int b2t_expand(b2t_handle* h, const b2t_expand_args* a, void* stream) {
    if (!h || !a) return B2T_ERR_ARG;
    if (a->n < 1 || a->n > h->cfg.n_games || !a->d_player || !a->d_K || !a->d_offsets || a->capacity < 0)
        return fail(h, B2T_ERR_ARG, "b2t_expand: need n>=1, d_player, d_K, d_offsets, capacity>=0%s");
    if (int r = check_obs(h, a->obs)) return r;
    cudaStream_t s = (cudaStream_t)stream;
    if (int r = b2t_enumerate(h, a->d_game_idx, a->n, a->d_player, a->d_K, nullptr, nullptr, 0, 0, stream)) return r;
    if (h->row_meta_cap < (size_t)a->capacity) {
        CK(cudaFree(h->row_meta));
        h->row_meta = nullptr; h->row_meta_cap = 0;
        CK(cudaMalloc((void**)&h->row_meta, ((size_t)a->capacity + 1) * sizeof(int4)));
        h->row_meta_cap = (size_t)a->capacity;
    }
    if (int r = scan_offsets(h, a->d_K, a->n, a->d_offsets, s)) return r;
    fill_row_meta<<<(a->n + 255) / 256, 256, 0, s>>>(a->d_offsets, a->n, a->d_game_idx, a->d_player, h->row_meta, a->d_board, a->capacity);
    CK(cudaGetLastError());
    KParams p; base_params(h, p);
    p.idx = a->d_game_idx; p.n = a->n; p.player = a->d_player;
    p.flags = KF_EXPAND | (a->finalize ? KF_FINALIZE : 0u);
    if (a->time_ms > 0) p.elapsed = a->time_ms;
    p.offsets = a->d_offsets; p.capacity = a->capacity; p.row_meta = h->row_meta;
    p.reward = a->d_reward; p.done = a->d_done; p.lines = a->d_lines;
    p.keys_out = a->d_keys; p.key_len_out = a->d_key_len; p.lmax_out = a->lmax;
    set_obs(p, a->obs);
    CK(launch_game(h, p, a->capacity, s));
    return B2T_OK;
}

int b2t_select(b2t_handle* h, const float* d_scores, const int32_t* d_offsets, int32_t n, int32_t capacity, int32_t* d_placement, void* stream) {
    if (!h || !d_scores || !d_offsets || !d_placement || n < 0 || capacity < 0) return fail(h, B2T_ERR_ARG, "b2t_select: NULL argument, n < 0 or capacity < 0%s");
    CK(cudaSetDevice(h->cfg.device));
    if (n) select_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_scores, d_offsets, n, capacity, d_placement);
    CK(cudaGetLastError());
    return B2T_OK;
}

int b2t_snapshot(b2t_handle* h, int32_t slot, void* stream) {
    if (!h || slot < 0 || slot >= B2T_SNAPSHOT_SLOTS) return fail(h, B2T_ERR_ARG, "b2t_snapshot: bad slot%s");
    CK(cudaSetDevice(h->cfg.device));
    if (!h->snap[slot]) CK(cudaMalloc((void**)&h->snap[slot], h->state_vecs * sizeof(uint4)));
    CK(cudaMemcpyAsync(h->snap[slot], h->state, h->state_vecs * sizeof(uint4), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (h->tiles) {
        if (!h->snap_tiles[slot]) CK(cudaMalloc((void**)&h->snap_tiles[slot], h->tile_words_total * sizeof(uint32_t)));
        CK(cudaMemcpyAsync(h->snap_tiles[slot], h->tiles, h->tile_words_total * sizeof(uint32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    }
    return B2T_OK;
}
int b2t_restore(b2t_handle* h, int32_t slot, void* stream) {
    if (!h || slot < 0 || slot >= B2T_SNAPSHOT_SLOTS || !h->snap[slot]) return fail(h, B2T_ERR_ARG, "b2t_restore: bad or empty slot%s");
    CK(cudaSetDevice(h->cfg.device));
    CK(cudaMemcpyAsync(h->state, h->snap[slot], h->state_vecs * sizeof(uint4), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (h->tiles) CK(cudaMemcpyAsync(h->tiles, h->snap_tiles[slot], h->tile_words_total * sizeof(uint32_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return B2T_OK;
}

int b2t_export_state(b2t_handle* h, int32_t game, void* h_buf, void* stream) {
    if (!h || !h_buf || game < 0 || game >= h->cfg.n_games) return fail(h, B2T_ERR_ARG, "b2t_export_state: bad argument%s");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    gather_game_kernel<<<(h->nv + 63) / 64, 64, 0, s>>>(h->state, game, h->nv, h->xfer);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h_buf, h->xfer, (size_t)h->nv * 16, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B2T_OK;
}
int b2t_import_state(b2t_handle* h, int32_t game, const void* h_buf, void* stream) {
    if (!h || !h_buf || game < 0 || game >= h->cfg.n_games) return fail(h, B2T_ERR_ARG, "b2t_import_state: bad argument%s");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    CK(cudaMemcpyAsync(h->xfer, h_buf, (size_t)h->nv * 16, cudaMemcpyHostToDevice, s));
    scatter_game_kernel<<<(h->nv + 63) / 64, 64, 0, s>>>(h->state, game, h->nv, h->xfer);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(s));
    return B2T_OK;
}

static int fetch_tiles(b2t_handle* h, int32_t game, std::vector<uint32_t>& planes) {
    planes.assign((size_t)h->ops->tile_words, 0u);
    gather_tiles_kernel<<<(h->ops->tile_words + 63) / 64, 64>>>(h->tiles, game, h->ops->tile_words, h->xfer_tiles);
    CK(cudaGetLastError());
    CK(cudaMemcpy(planes.data(), h->xfer_tiles, planes.size() * 4, cudaMemcpyDeviceToHost));
    return B2T_OK;
}
int b2t_get_game(b2t_handle* h, int32_t game, b2t_game_raw* out, uint8_t* h_field) {
    if (!h || !out || game < 0 || game >= h->cfg.n_games) return fail(h, B2T_ERR_ARG, "b2t_get_game: bad argument%s");
    if (h->multi()) return fail(h, B2T_ERR_UNSUPPORTED, "b2t_get_game: b2t_game_raw describes two players%s");
    std::vector<uint32_t> words((size_t)h->ops->NV * 4), planes;
    if (int r = b2t_export_state(h, game, words.data(), nullptr)) return r;
    if (h->tiles && h_field)
        if (int r = fetch_tiles(h, game, planes)) return r;
    h->ops->decode(words.data(), out, h_field, planes.empty() ? nullptr : planes.data());
    return B2T_OK;
}
int b2t_set_game(b2t_handle* h, int32_t game, const b2t_game_raw* in, const uint8_t* h_field) {
    if (!h || !in || game < 0 || game >= h->cfg.n_games) return fail(h, B2T_ERR_ARG, "b2t_set_game: bad argument%s");
    if (h->multi()) return fail(h, B2T_ERR_UNSUPPORTED, "b2t_set_game: b2t_game_raw describes two players%s");
    for (int p = 0; p < 2; p++)
        if (in->player[p].piece_kind < 0 || in->player[p].piece_kind > 7 || in->player[p].nextpiece < 0 || in->player[p].nextpiece > 7)
            return fail(h, B2T_ERR_ARG, "b2t_set_game: piece_kind / nextpiece out of range%s");
    std::vector<uint32_t> words((size_t)h->ops->NV * 4), planes;
    if (int r = b2t_export_state(h, game, words.data(), nullptr)) return r;
    if (h->tiles && h_field) planes.assign((size_t)h->ops->tile_words, 0u);
    h->ops->encode(words.data(), in, h_field, planes.empty() ? nullptr : planes.data());
    if (!planes.empty()) {
        CK(cudaMemcpy(h->xfer_tiles, planes.data(), planes.size() * 4, cudaMemcpyHostToDevice));
        scatter_tiles_kernel<<<(h->ops->tile_words + 63) / 64, 64>>>(h->tiles, game, h->ops->tile_words, h->xfer_tiles);
        CK(cudaGetLastError());
    }
    return b2t_import_state(h, game, words.data(), nullptr);
}
int b2t_copy_game(b2t_handle* to, int32_t dst, b2t_handle* from, int32_t src, void* stream) {
    b2t_handle* h = to;
    if (!to || !from || dst < 0 || dst >= to->cfg.n_games || src < 0 || src >= from->cfg.n_games)
        return fail(h, B2T_ERR_ARG, "b2t_copy_game: bad argument%s");
    if (to->ops != from->ops || to->cfg.device != from->cfg.device || to->np != from->np)
        return fail(h, B2T_ERR_ARG, "b2t_copy_game: handles differ in field size, player count or device%s");
    CK(cudaSetDevice(h->cfg.device));
    copy_game_kernel<<<(h->nv + 63) / 64, 64, 0, (cudaStream_t)stream>>>(to->state, dst, from->state, src, h->nv);
    if (to->tiles)   // colours travel with the game (a source without planes reads as all ones)
        copy_tiles_kernel<<<(h->ops->tile_words + 63) / 64, 64, 0, (cudaStream_t)stream>>>(to->tiles, dst, from->tiles, src, h->ops->tile_words);
    CK(cudaGetLastError());
    return B2T_OK;
}

static int poke(b2t_handle* h, int32_t game, int32_t player, uint32_t flag, int value, void* stream) {
    if (!h || game < 0 || game >= h->cfg.n_games || player < 0 || player >= h->np) return fail(h, B2T_ERR_ARG, "poke: bad game/player%s");
    CK(cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    int32_t* d_idx = reinterpret_cast<int32_t*>(h->d_total) + 4;
    CK(cudaMemcpyAsync(d_idx, &game, sizeof game, cudaMemcpyHostToDevice, s));
    KParams p; base_params(h, p);
    p.idx = d_idx; p.n = 1; p.flags = flag | KF_STORE; p.poke_player = player; p.poke_value = value;
    CK(launch_game(h, p, 1, s));
    CK(cudaStreamSynchronize(s));
    return B2T_OK;
}
int b2t_poke_dead(b2t_handle* h, int32_t game, int32_t player, int32_t dead, void* stream) { return poke(h, game, player, KF_POKE_DEAD, dead ? 1 : 0, stream); }
int b2t_clear_field(b2t_handle* h, int32_t game, int32_t player, void* stream) { return poke(h, game, player, KF_CLEAR_FIELD, 0, stream); }

int b2t_stats(b2t_handle* h, uint64_t* h_out, void* stream) {
    if (!h || !h_out) return B2T_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    unsigned long long tmp[STAT_SLICES * STAT_WORDS];
    CK(cudaMemcpyAsync(tmp, h->stats, sizeof tmp, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CK(cudaStreamSynchronize((cudaStream_t)stream));
    for (int j = 0; j < STAT_WORDS; j++) {
        h_out[j] = 0;
        for (int s = 0; s < STAT_SLICES; s++) h_out[j] += tmp[s * STAT_WORDS + j];
    }
    return B2T_OK;
}

int b2t_host_alloc(size_t bytes, int32_t write_combined, void** out) {
    if (!out || !bytes) return B2T_ERR_ARG;
    *out = nullptr;
    return cudaHostAlloc(out, bytes, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault) == cudaSuccess ? B2T_OK : B2T_ERR_CUDA;
}
int b2t_host_free(void* p) { return (!p || cudaFreeHost(p) == cudaSuccess) ? B2T_OK : B2T_ERR_CUDA; }

int b2t_debug_traffic(b2t_handle* h, int32_t n, const uint8_t* d_keys, int32_t lmax, const uint8_t* d_player, const uint8_t* d_key_len,
                      const uint32_t* d_seeds, int8_t* d_reward, uint8_t* d_done, uint8_t* d_lines, uint8_t* d_field, uint8_t* d_vec,
                      uint8_t* d_piece_idx, int32_t step_like, void* stream) {
    if (!h || n < 0 || n > h->cfg.n_games || !d_keys || !d_player || !d_key_len || !d_seeds || !d_reward || !d_done || !d_lines || !d_field ||
        !d_vec || !d_piece_idx) return fail(h, B2T_ERR_ARG, "b2t_debug_traffic: every array is required%s");
    CK(cudaSetDevice(h->cfg.device));
    CK(h->ops->launch_traffic(h->state, n, d_keys, lmax, d_player, d_key_len, d_seeds, d_reward, d_done, d_lines, d_field, d_vec, d_piece_idx,
                              step_like, (cudaStream_t)stream));
    return B2T_OK;
}

int b2t_debug_streams(b2t_handle* h, uint32_t seed16, int32_t n, uint8_t* h_pieces, uint8_t* h_holes) {
    if (!h || seed16 >= (uint32_t)RNG_SEEDS || n < 0) return B2T_ERR_ARG;
    CK(cudaSetDevice(h->cfg.device));
    const size_t np = (size_t)(n < h->cfg.piece_stream_len ? n : h->cfg.piece_stream_len), nh = (size_t)(n < h->cfg.hole_stream_len ? n : h->cfg.hole_stream_len);
    if (h_pieces) CK(cudaMemcpy(h_pieces, h->piece_tab + (size_t)seed16 * h->cfg.piece_stream_len, np, cudaMemcpyDeviceToHost));
    if (h_holes) CK(cudaMemcpy(h_holes, h->hole_tab + (size_t)seed16 * h->cfg.hole_stream_len, nh, cudaMemcpyDeviceToHost));
    return B2T_OK;
}

// ------------------------------------------------------------------------------------------ host-buffer entry points
struct In {   // stage one optional host array into the input arena
    size_t off = 0, bytes = 0;
    const void* src = nullptr;
};
static const void* dev_in(const Arena& a, const In& i) { return i.src ? a.d + i.off : nullptr; }
struct Out {
    size_t off = 0, bytes = 0;
    void* dst = nullptr;
};
static void* dev_out(const Arena& a, const Out& o) { return o.dst ? a.d + o.off : nullptr; }

// Pinned (cudaMallocHost / cudaHostRegister'ed, e.g. torch pin_memory) caller buffers are copied straight to / from the
// device arena; pageable ones go through the handle's pinned staging (one extra host memcpy).
constexpr size_t DIRECT_COPY_MIN = 1 << 16;   // smaller arrays go through the handle's pinned staging without asking the driver
// device address of a pinned / registered host range, or NULL: the first AND the last byte must resolve to host memory of one
// mapping (a buffer that is only partly registered does not pass)
static void* mapped_range(const void* p, size_t bytes) {
    if (!p) return nullptr;
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a0.type != cudaMemoryTypeHost || !a0.devicePointer) return nullptr;
    if (bytes > 1) {
        if (cudaPointerGetAttributes(&a1, (const unsigned char*)p + bytes - 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        if (a1.type != cudaMemoryTypeHost || (unsigned char*)a1.devicePointer != (unsigned char*)a0.devicePointer + bytes - 1) return nullptr;
    }
    return a0.devicePointer;
}
static bool is_pinned(const void* p, size_t bytes) { return mapped_range(p, bytes) != nullptr; }
static int stage_in(b2t_handle* h, std::vector<In*> ins, cudaStream_t s) {
    size_t total = 0;
    for (auto i : ins) if (i->src) total += (i->bytes + 255) & ~(size_t)255;
    CK(h->in.reserve(total + 256));
    for (auto i : ins) {
        if (!i->src || !i->bytes) { if (i->src) i->off = h->in.take(1); continue; }
        i->off = h->in.take(i->bytes);
        const void* from = i->src;
        if (i->bytes < DIRECT_COPY_MIN || !is_pinned(from, i->bytes)) { memcpy(h->in.h + i->off, i->src, i->bytes); from = h->in.h + i->off; }
        CK(cudaMemcpyAsync(h->in.d + i->off, from, i->bytes, cudaMemcpyHostToDevice, s));
    }
    return B2T_OK;
}
static int plan_out(b2t_handle* h, std::vector<Out*> outs) {
    size_t total = 0;
    for (auto o : outs) if (o->dst) total += (o->bytes + 255) & ~(size_t)255;
    CK(h->out.reserve(total + 256));
    for (auto o : outs) if (o->dst) o->off = h->out.take(o->bytes ? o->bytes : 1);
    return B2T_OK;
}
static int fetch_out(b2t_handle* h, std::vector<Out*> outs, cudaStream_t s) {
    std::vector<Out*> staged;
    for (auto o : outs) {
        if (!o->dst || !o->bytes) continue;
        if (o->bytes >= DIRECT_COPY_MIN && is_pinned(o->dst, o->bytes)) CK(cudaMemcpyAsync(o->dst, h->out.d + o->off, o->bytes, cudaMemcpyDeviceToHost, s));
        else { CK(cudaMemcpyAsync(h->out.h + o->off, h->out.d + o->off, o->bytes, cudaMemcpyDeviceToHost, s)); staged.push_back(o); }
    }
    CK(cudaStreamSynchronize(s));
    for (auto o : staged) memcpy(o->dst, h->out.h + o->off, o->bytes);
    return B2T_OK;
}

int b2t_export_states(b2t_handle* h, const int32_t* h_idx, int32_t n, void* h_buf) {
    if (!h || !h_buf || n < 0 || n > h->cfg.n_games) return fail(h, B2T_ERR_ARG, "b2t_export_states: bad argument%s");
    if (!n) return B2T_OK;
    CK_IDS("b2t_export_states");
    CK(cudaSetDevice(h->cfg.device));
    const size_t bytes = (size_t)n * h->nv * 16;
    In idx; idx.src = h_idx; idx.bytes = (size_t)n * 4;
    if (int r = stage_in(h, {&idx}, 0)) return r;
    Out buf; buf.dst = h_buf; buf.bytes = bytes;
    if (int r = plan_out(h, {&buf})) return r;
    const long long t = (long long)n * h->nv;
    gather_games_kernel<<<(unsigned)((t + 255) / 256), 256>>>(h->state, h->cfg.n_games, (const int32_t*)dev_in(h->in, idx), n, h->nv, (uint4*)dev_out(h->out, buf));
    CK(cudaGetLastError());
    return fetch_out(h, {&buf}, 0);
}
int b2t_import_states(b2t_handle* h, const int32_t* h_idx, int32_t n, const void* h_buf) {
    if (!h || !h_buf || n < 0 || n > h->cfg.n_games) return fail(h, B2T_ERR_ARG, "b2t_import_states: bad argument%s");
    if (!n) return B2T_OK;
    CK_IDS("b2t_import_states");
    CK(cudaSetDevice(h->cfg.device));
    In idx, buf; idx.src = h_idx; idx.bytes = (size_t)n * 4; buf.src = h_buf; buf.bytes = (size_t)n * h->nv * 16;
    if (int r = stage_in(h, {&idx, &buf}, 0)) return r;
    const long long t = (long long)n * h->nv;
    scatter_games_kernel<<<(unsigned)((t + 255) / 256), 256>>>(h->state, h->cfg.n_games, (const int32_t*)dev_in(h->in, idx), n, h->nv, (const uint4*)dev_in(h->in, buf));
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(0));
    return B2T_OK;
}

int b2t_reset_host(b2t_handle* h, const int32_t* h_idx, int32_t n, const uint32_t* h_seeds) {
    if (!h) return B2T_ERR_ARG;
    if (n < 0 || n > h->cfg.n_games || !h_seeds) return fail(h, B2T_ERR_ARG, "b2t_reset_host: bad n or NULL seeds%s");
    CK_IDS("b2t_reset_host");
    CK(cudaSetDevice(h->cfg.device));
    In idx, seeds;
    idx.src = h_idx; idx.bytes = (size_t)n * 4; seeds.src = h_seeds; seeds.bytes = (size_t)n * 4;
    if (int r = stage_in(h, {&idx, &seeds}, 0)) return r;
    if (int r = b2t_reset(h, (const int32_t*)dev_in(h->in, idx), n, (const uint32_t*)dev_in(h->in, seeds), nullptr)) return r;
    CK(cudaStreamSynchronize(0));
    return B2T_OK;
}

int b2t_step_host(b2t_handle* h, const int32_t* h_idx, int32_t n, const uint8_t* h_player, const uint8_t* h_keys, const uint8_t* h_key_len,
                  int32_t lmax, const int32_t* h_placement, int32_t finalize, int32_t time_ms, const uint32_t* h_reset_seeds,
                  int8_t* h_reward, uint8_t* h_done, uint8_t* h_lines, uint8_t* h_field, uint8_t* h_vec, uint8_t* h_piece_idx,
                  int32_t* h_scalars, int32_t key_format, int32_t auto_reset) {
    if (!h) return B2T_ERR_ARG;
    if (n < 0 || n > h->cfg.n_games) return fail(h, B2T_ERR_ARG, "b2t_step_host: bad n%s");
    CK_IDS("b2t_step_host");
    CK(cudaSetDevice(h->cfg.device));
    const size_t N = (size_t)n, HW = (size_t)h->ops->H * h->ops->W, P = (size_t)h->np;
    In idx, pl, keys, klen, plc, rs;
    idx.src = h_idx; idx.bytes = N * 4; pl.src = h_player; pl.bytes = N;
    const size_t key_row = (size_t)(lmax > 0 ? lmax : 0) >> (key_format == B2T_KEYS_PACKED4 ? 1 : 0);
    keys.src = h_keys; keys.bytes = N * key_row; klen.src = h_key_len; klen.bytes = N;
    plc.src = h_placement; plc.bytes = N * 4; rs.src = h_reset_seeds; rs.bytes = N * 4;
    Out rew, done, lines, field, vec, pidx, scal;
    rew.dst = h_reward; rew.bytes = N; done.dst = h_done; done.bytes = N; lines.dst = h_lines; lines.bytes = P * N;
    field.dst = h_field; field.bytes = P * HW * N; vec.dst = h_vec; vec.bytes = P * B2T_VEC_LEN * N;
    pidx.dst = h_piece_idx; pidx.bytes = P * N; scal.dst = h_scalars; scal.bytes = 32 * P * N;
    {
        // Large batch, every buffer pinned, no observation tensors asked for: the kernel reads the keys and writes reward / done /
        // lines straight through the caller's pinned pages (they are device-addressable under unified addressing), so the
        // PCIe traffic overlaps the compute inside ONE launch instead of being copies before and after it.
        static const bool zero_copy = !(getenv("B2T_ZERO_COPY") && atoi(getenv("B2T_ZERO_COPY")) == 0);
        // the kernel loads ids / placements / seeds as 32-bit words and stores the line counts as 16-bit pairs: a host pointer that
        // is not naturally aligned for that (a sliced array) takes the staged path instead of faulting on the device
        auto mapped = [](const void* p, size_t bytes, size_t align) -> void* {
            if (!p || (reinterpret_cast<uintptr_t>(p) & (align - 1))) return nullptr;
            void* d = mapped_range(p, bytes);
            return (d && !(reinterpret_cast<uintptr_t>(d) & (align - 1))) ? d : nullptr;
        };
        if (zero_copy && n >= PIPE_MIN_ROWS && !h_field && !h_vec && !h_piece_idx && !h_scalars) {
            const void* ins_h[6] = {h_idx, h_player, h_keys, h_key_len, h_placement, h_reset_seeds};
            const size_t ins_b[6] = {idx.bytes, pl.bytes, keys.bytes, klen.bytes, plc.bytes, rs.bytes}, ins_a[6] = {4, 1, 1, 1, 4, 4};
            void* outs_h[3] = {h_reward, h_done, h_lines};
            const size_t outs_b[3] = {rew.bytes, done.bytes, lines.bytes}, outs_a[3] = {1, 1, 2};
            void* ins_d[6];
            void* outs_d[3];
            bool ok = true;
            for (int k = 0; k < 6; k++) { ins_d[k] = mapped(ins_h[k], ins_b[k], ins_a[k]); ok = ok && (!ins_h[k] || ins_d[k]); }
            for (int k = 0; k < 3; k++) { outs_d[k] = mapped(outs_h[k], outs_b[k], outs_a[k]); ok = ok && (!outs_h[k] || outs_d[k]); }
            if (ok) {
                b2t_step_args a;
                memset(&a, 0, sizeof a);
                a.d_game_idx = (const int32_t*)ins_d[0]; a.n = n; a.d_player = (const uint8_t*)ins_d[1];
                a.d_keys = (const uint8_t*)ins_d[2]; a.d_key_len = (const uint8_t*)ins_d[3]; a.lmax = lmax;
                a.d_placement = (const int32_t*)ins_d[4]; a.finalize = finalize; a.time_ms = time_ms;
                a.d_reset_seeds = (const uint32_t*)ins_d[5]; a.key_format = key_format; a.auto_reset = auto_reset;
                a.d_reward = (int8_t*)outs_d[0]; a.d_done = (uint8_t*)outs_d[1]; a.d_lines = (uint8_t*)outs_d[2];
                if (int r = step_launch(h, &a, 0, nullptr)) return r;
                CK(cudaStreamSynchronize(0));
                return B2T_OK;
            }
        }
    }
    if (n >= PIPE_MIN_ROWS) {
        // Large batch: slices of the rows go down, through the kernel and back up on two streams, so the PCIe copies of one
        // slice overlap the kernel of another (a step of 2^20 games is 23 MB in, 0.25 ms of kernel, 4 MB out).
        std::vector<In*> ins = {&idx, &pl, &keys, &klen, &plc, &rs};
        std::vector<Out*> outs = {&rew, &done, &lines, &field, &vec, &pidx, &scal};
        const size_t in_row[6] = {4, 1, key_row, 1, 4, 4};
        const size_t out_row[7] = {1, 1, P, P * HW, P * B2T_VEC_LEN, P, 32 * P};
        size_t total = 0;
        for (auto i : ins) if (i->src) total += (i->bytes + 255) & ~(size_t)255;
        CK(h->in.reserve(total + 256));
        for (auto i : ins) if (i->src) i->off = h->in.take(i->bytes ? i->bytes : 1);
        if (int r = plan_out(h, outs)) return r;
        for (auto& st : h->pipe) if (!st) CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        CK(cudaStreamSynchronize(0));               // earlier work of this handle went to the default stream
        StreamDrain drain{h->pipe[0], h->pipe[1]};
        bool pinned_in[6], pinned_out[7];
        for (int k = 0; k < 6; k++) pinned_in[k] = ins[k]->src && is_pinned(ins[k]->src, ins[k]->bytes);
        for (int k = 0; k < 7; k++) pinned_out[k] = outs[k]->dst && is_pinned(outs[k]->dst, outs[k]->bytes);
        static const int slices_env = getenv("B2T_PIPE_SLICES") ? atoi(getenv("B2T_PIPE_SLICES")) : 0;   // tuning knob
        const int slices = slices_env > 0 ? slices_env : PIPE_SLICES;
        const int per = (((n + slices - 1) / slices) + 1023) & ~1023;      // slice starts stay 16-byte aligned in every array
        for (int c = 0, lo = 0; lo < n; c++, lo += per) {
            const int cnt = (n - lo) < per ? (n - lo) : per;
            cudaStream_t st = h->pipe[c & 1];
            for (int k = 0; k < 6; k++) {
                In* i = ins[k];
                if (!i->src || !in_row[k]) continue;
                const size_t o = (size_t)lo * in_row[k], b = (size_t)cnt * in_row[k];
                const unsigned char* from = (const unsigned char*)i->src + o;
                if (!pinned_in[k]) { memcpy(h->in.h + i->off + o, from, b); from = h->in.h + i->off + o; }
                CK(cudaMemcpyAsync(h->in.d + i->off + o, from, b, cudaMemcpyHostToDevice, st));
            }
            b2t_step_args a;
            memset(&a, 0, sizeof a);
            auto din = [&](const In& i, int k) -> const void* { return i.src ? h->in.d + i.off + (size_t)lo * in_row[k] : nullptr; };
            auto dout = [&](const Out& o, int k) -> void* { return o.dst ? h->out.d + o.off + (size_t)lo * out_row[k] : nullptr; };
            a.d_game_idx = (const int32_t*)din(idx, 0); a.n = cnt; a.d_player = (const uint8_t*)din(pl, 1);
            a.d_keys = (const uint8_t*)din(keys, 2); a.d_key_len = (const uint8_t*)din(klen, 3); a.lmax = lmax;
            a.d_placement = (const int32_t*)din(plc, 4); a.finalize = finalize; a.time_ms = time_ms;
            a.d_reset_seeds = (const uint32_t*)din(rs, 5); a.key_format = key_format; a.auto_reset = auto_reset;
            a.d_reward = (int8_t*)dout(rew, 0); a.d_done = (uint8_t*)dout(done, 1); a.d_lines = (uint8_t*)dout(lines, 2);
            a.obs.d_field = dout(field, 3); a.obs.d_vec = dout(vec, 4);
            a.obs.d_piece_idx = (uint8_t*)dout(pidx, 5); a.obs.d_scalars = (int32_t*)dout(scal, 6);
            if (int r = step_launch(h, &a, h_idx ? 0 : lo, st)) return r;
            for (int k = 0; k < 7; k++) {
                Out* o = outs[k];
                if (!o->dst) continue;
                const size_t off = (size_t)lo * out_row[k], b = (size_t)cnt * out_row[k];
                void* to = pinned_out[k] ? (unsigned char*)o->dst + off : h->out.h + o->off + off;
                CK(cudaMemcpyAsync(to, h->out.d + o->off + off, b, cudaMemcpyDeviceToHost, st));
            }
        }
        CK(cudaStreamSynchronize(h->pipe[0]));
        CK(cudaStreamSynchronize(h->pipe[1]));
        for (int k = 0; k < 7; k++)
            if (outs[k]->dst && !pinned_out[k]) memcpy(outs[k]->dst, h->out.h + outs[k]->off, outs[k]->bytes);
        return B2T_OK;
    }
    if (int r = stage_in(h, {&idx, &pl, &keys, &klen, &plc, &rs}, 0)) return r;
    rew.dst = h_reward; rew.bytes = N; done.dst = h_done; done.bytes = N; lines.dst = h_lines; lines.bytes = P * N;
    field.dst = h_field; field.bytes = P * HW * N; vec.dst = h_vec; vec.bytes = P * B2T_VEC_LEN * N;
    pidx.dst = h_piece_idx; pidx.bytes = P * N; scal.dst = h_scalars; scal.bytes = 32 * P * N;
    if (int r = plan_out(h, {&rew, &done, &lines, &field, &vec, &pidx, &scal})) return r;
    b2t_step_args a;
    memset(&a, 0, sizeof a);
    a.d_game_idx = (const int32_t*)dev_in(h->in, idx); a.n = n; a.d_player = (const uint8_t*)dev_in(h->in, pl);
    a.d_keys = (const uint8_t*)dev_in(h->in, keys); a.d_key_len = (const uint8_t*)dev_in(h->in, klen); a.lmax = lmax;
    a.d_placement = (const int32_t*)dev_in(h->in, plc); a.finalize = finalize; a.time_ms = time_ms;
    a.d_reset_seeds = (const uint32_t*)dev_in(h->in, rs); a.key_format = key_format; a.auto_reset = auto_reset;
    a.d_reward = (int8_t*)dev_out(h->out, rew); a.d_done = (uint8_t*)dev_out(h->out, done); a.d_lines = (uint8_t*)dev_out(h->out, lines);
    a.obs.d_field = (uint8_t*)dev_out(h->out, field); a.obs.d_vec = (uint8_t*)dev_out(h->out, vec);
    a.obs.d_piece_idx = (uint8_t*)dev_out(h->out, pidx); a.obs.d_scalars = (int32_t*)dev_out(h->out, scal);
    if (int r = b2t_step(h, &a, nullptr)) return r;
    return fetch_out(h, {&rew, &done, &lines, &field, &vec, &pidx, &scal}, 0);
}

int b2t_observe_host(b2t_handle* h, const int32_t* h_idx, int32_t n, uint8_t* h_field, uint8_t* h_vec, uint8_t* h_piece_idx, int32_t* h_scalars) {
    if (!h) return B2T_ERR_ARG;
    if (n < 0 || n > h->cfg.n_games) return fail(h, B2T_ERR_ARG, "b2t_observe_host: bad n%s");
    CK_IDS("b2t_observe_host");
    CK(cudaSetDevice(h->cfg.device));
    const size_t N = (size_t)n, HW = (size_t)h->ops->H * h->ops->W, P = (size_t)h->np;
    In idx; idx.src = h_idx; idx.bytes = N * 4;
    if (int r = stage_in(h, {&idx}, 0)) return r;
    Out field, vec, pidx, scal;
    field.dst = h_field; field.bytes = P * HW * N; vec.dst = h_vec; vec.bytes = P * B2T_VEC_LEN * N;
    pidx.dst = h_piece_idx; pidx.bytes = P * N; scal.dst = h_scalars; scal.bytes = 32 * P * N;
    if (int r = plan_out(h, {&field, &vec, &pidx, &scal})) return r;
    b2t_obs o;
    memset(&o, 0, sizeof o);
    o.d_field = (uint8_t*)dev_out(h->out, field); o.d_vec = (uint8_t*)dev_out(h->out, vec);
    o.d_piece_idx = (uint8_t*)dev_out(h->out, pidx); o.d_scalars = (int32_t*)dev_out(h->out, scal);
    if (int r = b2t_observe(h, (const int32_t*)dev_in(h->in, idx), n, &o, nullptr)) return r;
    return fetch_out(h, {&field, &vec, &pidx, &scal}, 0);
}

int b2t_enumerate_host(b2t_handle* h, const int32_t* h_idx, int32_t n, const uint8_t* h_player, int32_t* h_K, uint8_t* h_keys,
                       uint8_t* h_key_len, int32_t kmax, int32_t lmax) {
    if (!h) return B2T_ERR_ARG;
    if (n < 0 || n > h->cfg.n_games || !h_player) return fail(h, B2T_ERR_ARG, "b2t_enumerate_host: bad n or NULL player%s");
    CK_IDS("b2t_enumerate_host");
    CK(cudaSetDevice(h->cfg.device));
    const size_t N = (size_t)n;
    In idx, pl; idx.src = h_idx; idx.bytes = N * 4; pl.src = h_player; pl.bytes = N;
    if (int r = stage_in(h, {&idx, &pl}, 0)) return r;
    Out K, keys, klen;
    K.dst = h_K; K.bytes = N * 4; keys.dst = h_keys; keys.bytes = N * (size_t)kmax * (size_t)lmax; klen.dst = h_key_len; klen.bytes = N * (size_t)kmax;
    if (int r = plan_out(h, {&K, &keys, &klen})) return r;
    if (int r = b2t_enumerate(h, (const int32_t*)dev_in(h->in, idx), n, (const uint8_t*)dev_in(h->in, pl), (int32_t*)dev_out(h->out, K),
                              (uint8_t*)dev_out(h->out, keys), (uint8_t*)dev_out(h->out, klen), kmax, lmax, nullptr)) return r;
    return fetch_out(h, {&K, &keys, &klen}, 0);
}

int b2t_expand_host(b2t_handle* h, const int32_t* h_idx, int32_t n, const uint8_t* h_player, int32_t finalize, int32_t time_ms,
                    int32_t capacity, int32_t* h_K, int32_t* h_offsets, int32_t* h_total, uint8_t* h_keys, uint8_t* h_key_len, int32_t lmax,
                    int8_t* h_reward, uint8_t* h_done, uint8_t* h_lines, uint8_t* h_field, uint8_t* h_vec, uint8_t* h_piece_idx,
                    int32_t* h_scalars) {
    if (!h) return B2T_ERR_ARG;
    if (n < 1 || n > h->cfg.n_games || !h_player || capacity < 0) return fail(h, B2T_ERR_ARG, "b2t_expand_host: bad n/capacity or NULL player%s");
    CK_IDS("b2t_expand_host");
    CK(cudaSetDevice(h->cfg.device));
    const size_t N = (size_t)n, C = (size_t)capacity, HW = (size_t)h->ops->H * h->ops->W, P = (size_t)h->np;
    In idx, pl; idx.src = h_idx; idx.bytes = N * 4; pl.src = h_player; pl.bytes = N;
    if (int r = stage_in(h, {&idx, &pl}, 0)) return r;
    Out K, offs, keys, klen, rew, done, lines, field, vec, pidx, scal;
    std::vector<int32_t> k_tmp, o_tmp;
    if (!h_K) { k_tmp.resize(N); h_K = k_tmp.data(); }
    if (!h_offsets) { o_tmp.resize(N + 1); h_offsets = o_tmp.data(); }
    K.dst = h_K; K.bytes = N * 4; offs.dst = h_offsets; offs.bytes = (N + 1) * 4;
    keys.dst = h_keys; keys.bytes = C * (size_t)(lmax > 0 ? lmax : 0); klen.dst = h_key_len; klen.bytes = C;
    rew.dst = h_reward; rew.bytes = C; done.dst = h_done; done.bytes = C; lines.dst = h_lines; lines.bytes = P * C;
    field.dst = h_field; field.bytes = P * HW * C; vec.dst = h_vec; vec.bytes = P * B2T_VEC_LEN * C; pidx.dst = h_piece_idx; pidx.bytes = P * C;
    scal.dst = h_scalars; scal.bytes = 32 * P * C;
    if (int r = plan_out(h, {&K, &offs, &keys, &klen, &rew, &done, &lines, &field, &vec, &pidx, &scal})) return r;
    b2t_expand_args a;
    memset(&a, 0, sizeof a);
    a.d_game_idx = (const int32_t*)dev_in(h->in, idx); a.n = n; a.d_player = (const uint8_t*)dev_in(h->in, pl);
    a.finalize = finalize; a.time_ms = time_ms; a.capacity = capacity;
    a.d_K = (int32_t*)dev_out(h->out, K); a.d_offsets = (int32_t*)dev_out(h->out, offs);
    a.d_keys = (uint8_t*)dev_out(h->out, keys); a.d_key_len = (uint8_t*)dev_out(h->out, klen); a.lmax = lmax;
    a.d_reward = (int8_t*)dev_out(h->out, rew); a.d_done = (uint8_t*)dev_out(h->out, done); a.d_lines = (uint8_t*)dev_out(h->out, lines);
    a.obs.d_field = (uint8_t*)dev_out(h->out, field); a.obs.d_vec = (uint8_t*)dev_out(h->out, vec); a.obs.d_piece_idx = (uint8_t*)dev_out(h->out, pidx);
    a.obs.d_scalars = (int32_t*)dev_out(h->out, scal);
    if (int r = b2t_expand(h, &a, nullptr)) return r;
    if (int r = fetch_out(h, {&K, &offs, &keys, &klen, &rew, &done, &lines, &field, &vec, &pidx, &scal}, 0)) return r;
    if (h_total) *h_total = h_offsets[n];
    return B2T_OK;
}

}  // extern "C"
