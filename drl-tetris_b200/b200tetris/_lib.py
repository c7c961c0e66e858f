"""ctypes binding of the C ABI in include/b200tetris.h (libb200tetris.so, built in-tree by drl-tetris_b200/build.py).

There is no Python or CPU fallback: if the shared library is missing or no CUDA device is present the import /
`b2t_create` fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B2T_LIB") or os.path.join(_HERE, "libb200tetris.so")   # B2T_LIB: experimental builds (build.py --out)

VEC_LEN = 19
U8, F16, BF16, F32, BITS = 0, 1, 2, 3, 4   # b2t_dtype
FLAG_TILES = 1                    # b2t_config.flags
KEYS_U8, KEYS_PACKED4 = 0, 1      # b2t_step_args.key_format
SNAPSHOT_SLOTS = 4

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_CAPACITY, ERR_STATE = 0, -1, -2, -3, -4, -5

u8p, i8p, i32p, u32p, u64p, vp = (C.POINTER(C.c_uint8), C.POINTER(C.c_int8), C.POINTER(C.c_int32), C.POINTER(C.c_uint32),
                                  C.POINTER(C.c_uint64), C.c_void_p)


class Config(C.Structure):
    _fields_ = [("n_games", C.c_int32), ("n_players", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("piece_map", C.c_uint8 * 7), ("bar_null_moves", C.c_uint8), ("time_ms_per_action", C.c_int32),
                ("piece_stream_len", C.c_int32), ("hole_stream_len", C.c_int32), ("device", C.c_int32), ("flags", C.c_uint32), ("seed", C.c_uint32)]


class Info(C.Structure):
    _fields_ = [("kmax", C.c_int32), ("lmax", C.c_int32), ("state_bytes", C.c_int32), ("obs_bytes", C.c_int32),
                ("tile_bytes", C.c_int32)]


class Obs(C.Structure):
    _fields_ = [("d_field", vp), ("d_vec", vp), ("d_piece_idx", vp), ("d_scalars", vp), ("perspective", C.c_int32), ("d_me", vp), ("dtype", C.c_int32)]


class StepArgs(C.Structure):
    _fields_ = [("d_game_idx", vp), ("n", C.c_int32), ("d_player", vp), ("d_keys", vp), ("d_key_len", vp), ("lmax", C.c_int32),
                ("d_placement", vp), ("finalize", C.c_int32), ("time_ms", C.c_int32), ("d_reset_seeds", vp),
                ("d_reward", vp), ("d_done", vp), ("d_lines", vp), ("obs", Obs), ("key_format", C.c_int32), ("auto_reset", C.c_int32)]


class ExpandArgs(C.Structure):
    _fields_ = [("d_game_idx", vp), ("n", C.c_int32), ("d_player", vp), ("finalize", C.c_int32), ("time_ms", C.c_int32),
                ("capacity", C.c_int32), ("d_K", vp), ("d_offsets", vp), ("d_board", vp), ("d_keys", vp), ("d_key_len", vp),
                ("lmax", C.c_int32), ("d_reward", vp), ("d_done", vp), ("d_lines", vp), ("obs", Obs)]


class PlayerRaw(C.Structure):   # b2t_player_raw
    _fields_ = [("piece", C.c_uint8 * 16)] + [(k, C.c_int32) for k in (
        "x", "y", "inc_lines", "combo_time", "combo_count", "nextpiece", "reward", "dead", "piece_kind", "rotation", "time_ms",
        "drop_delay", "drop_delay_time", "increase_drop_delay_time", "lockdown_time", "lockdown", "combo_start", "combo_time_total",
        "line_count", "min_remaining", "incoming_lines", "lines_pending", "garbage_n")] + [
        ("garbage_count", C.c_int32 * 6), ("garbage_delay", C.c_int32 * 6), ("piece_draws", C.c_int32), ("hole_draws", C.c_int32)]


class GameRaw(C.Structure):     # b2t_game_raw
    _fields_ = [("seed16", C.c_int32), ("round_over", C.c_int32), ("done", C.c_int32), ("last_winner", C.c_int32),
                ("overflow", C.c_int32), ("rounds_started", C.c_int32), ("player", PlayerRaw * 2)]


SYMBOLS = {
    "b2t_create": (C.c_int, [C.POINTER(Config), C.POINTER(vp)]),
    "b2t_destroy": (C.c_int, [vp]),
    "b2t_last_error": (C.c_char_p, [vp]),
    "b2t_get_info": (C.c_int, [vp, C.POINTER(Info)]),
    "b2t_reset": (C.c_int, [vp, vp, C.c_int32, vp, vp]),
    "b2t_step": (C.c_int, [vp, C.POINTER(StepArgs), vp]),
    "b2t_observe": (C.c_int, [vp, vp, C.c_int32, C.POINTER(Obs), vp]),
    "b2t_enumerate": (C.c_int, [vp, vp, C.c_int32, vp, vp, vp, vp, C.c_int32, C.c_int32, vp]),
    "b2t_expand": (C.c_int, [vp, C.POINTER(ExpandArgs), vp]),
    "b2t_select": (C.c_int, [vp, vp, vp, C.c_int32, C.c_int32, vp, vp]),
    "b2t_snapshot": (C.c_int, [vp, C.c_int32, vp]),
    "b2t_restore": (C.c_int, [vp, C.c_int32, vp]),
    "b2t_export_state": (C.c_int, [vp, C.c_int32, vp, vp]),
    "b2t_import_state": (C.c_int, [vp, C.c_int32, vp, vp]),
    "b2t_export_states": (C.c_int, [vp, i32p, C.c_int32, vp]),
    "b2t_import_states": (C.c_int, [vp, i32p, C.c_int32, vp]),
    "b2t_get_game": (C.c_int, [vp, C.c_int32, C.POINTER(GameRaw), u8p]),
    "b2t_set_game": (C.c_int, [vp, C.c_int32, C.POINTER(GameRaw), u8p]),
    "b2t_copy_game": (C.c_int, [vp, C.c_int32, vp, C.c_int32, vp]),
    "b2t_poke_dead": (C.c_int, [vp, C.c_int32, C.c_int32, C.c_int32, vp]),
    "b2t_clear_field": (C.c_int, [vp, C.c_int32, C.c_int32, vp]),
    "b2t_stats": (C.c_int, [vp, u64p, vp]),
    "b2t_reset_host": (C.c_int, [vp, i32p, C.c_int32, u32p]),
    "b2t_step_host": (C.c_int, [vp, i32p, C.c_int32, u8p, u8p, u8p, C.c_int32, i32p, C.c_int32, C.c_int32, u32p,
                                i8p, u8p, u8p, u8p, u8p, u8p, i32p, C.c_int32, C.c_int32]),
    "b2t_observe_host": (C.c_int, [vp, i32p, C.c_int32, u8p, u8p, u8p, i32p]),
    "b2t_enumerate_host": (C.c_int, [vp, i32p, C.c_int32, u8p, i32p, u8p, u8p, C.c_int32, C.c_int32]),
    "b2t_expand_host": (C.c_int, [vp, i32p, C.c_int32, u8p, C.c_int32, C.c_int32, C.c_int32, i32p, i32p, i32p,
                                  u8p, u8p, C.c_int32, i8p, u8p, u8p, u8p, u8p, u8p, i32p]),
    "b2t_host_alloc": (C.c_int, [C.c_size_t, C.c_int32, C.POINTER(vp)]),
    "b2t_host_free": (C.c_int, [vp]),
    "b2t_debug_traffic": (C.c_int, [vp, C.c_int32, vp, C.c_int32, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_int32, vp]),
    "b2t_debug_streams": (C.c_int, [vp, C.c_uint32, C.c_int32, u8p, u8p]),
}

_lib = None


def load():
    """dlopen the in-tree library and type every symbol the header declares."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH) and not os.environ.get("B2T_LIB"):
            # a fresh checkout: compile the CUDA library in-tree (nvcc cross-compiles sm_100a without a GPU); never a CPU fallback
            try:
                import importlib.util
                spec = importlib.util.spec_from_file_location("b2t_build", os.path.join(os.path.dirname(_HERE), "build.py"))
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                mod.build(verbose=False)
            except Exception as e:  # noqa: BLE001
                raise ImportError("b200tetris: %s is missing and could not be built (%s) — run `python drl-tetris_b200/build.py` "
                                  "(nvcc, sm_100a).  There is no CPU fallback." % (LIB_PATH, e))
        if not os.path.exists(LIB_PATH):
            raise ImportError("b200tetris: %s is missing.  There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)   # AttributeError here = header/library mismatch
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


class B2TError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200tetris error %d: %s" % (code, msg))
        self.code = code
