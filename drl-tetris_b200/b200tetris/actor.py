"""ActorLoop: one SIXten-style ply for every game of an Engine, entirely on the device:

    enumerate + expand (all placements of the side to move, observation tensors in (me, opponent) order and in the network's
    dtype)  ->  the caller's scoring function on those tensors  ->  select (arg-max per board)  ->  step by placement index
    with fused reset (engine-chosen seeds)  ->  observation of the new states

The reference does this with a Python loop over envs, two deep copies of the backend per simulated placement and a host
round trip per call (drl_tetris/worker.py:91-118, agents/sventon_agent/sventon_agent.py:56-98, tetris_environment.py:87-100).
Here every stage is an asynchronous C-ABI call on torch's current stream, so after one warm-up ply the whole chain is captured
in a CUDA graph and a ply is a single graph launch: no allocation, no synchronisation, no host data.  torch is only memory,
streams and the graph container."""


class ActorLoop:
    def __init__(self, engine, score_fn, rows_per_game=40, dtype=None, perspective=True, use_graph=True):
        """score_fn(field[rows,2,H,W], vec[rows,2,19], piece_idx[rows,2], scores[rows]) writes one float32 score per simulated
        placement (rows past offsets[-1] hold stale data; their scores are never read).  rows_per_game sizes the placement
        buffers (mean list length is ~25 at 22x10, so small vectors should pass rows_per_game=engine.kmax).  A ply whose placements
        outgrow the buffers stays memory-safe — the rows that did not fit are not simulated and select never reads past the buffers,
        so the boards at the end of the vector choose among the rows they kept, or placement 0 — and is reported by `overflowed()`."""
        import torch
        self.torch, self.e, self.score_fn = torch, engine, score_fn
        n, H, W = engine.n, engine.H, engine.W
        dev = torch.device("cuda", engine.device)
        dt = dtype or torch.uint8
        self.n, self.cap, self.perspective = n, int(n * rows_per_game), bool(perspective)
        self.player = torch.zeros(n, dtype=torch.uint8, device=dev)
        self.K = torch.empty(n, dtype=torch.int32, device=dev)
        self.offsets = torch.empty(n + 1, dtype=torch.int32, device=dev)
        self.p_field = torch.empty((self.cap, 2, H, W), dtype=dt, device=dev)
        self.p_vec = torch.empty((self.cap, 2, 19), dtype=dt, device=dev)
        self.p_piece = torch.empty((self.cap, 2), dtype=torch.uint8, device=dev)
        self.scores = torch.zeros(self.cap, dtype=torch.float32, device=dev)
        self.choice = torch.empty(n, dtype=torch.int32, device=dev)
        self.reward = torch.empty(n, dtype=torch.int8, device=dev)
        self.done = torch.empty(n, dtype=torch.uint8, device=dev)
        self.field = torch.empty((n, 2, H, W), dtype=dt, device=dev)
        self.vec = torch.empty((n, 2, 19), dtype=dt, device=dev)
        self.piece = torch.empty((n, 2), dtype=torch.uint8, device=dev)
        self.plies = 0
        self._graph = None
        self._use_graph = bool(use_graph)

    def _ply(self):
        e = self.e
        e.expand_dev(self.player, self.K, self.offsets, self.cap, field=self.p_field, vec=self.p_vec, piece_idx=self.p_piece,
                     perspective=self.perspective)
        self.score_fn(self.p_field, self.p_vec, self.p_piece, self.scores)
        e.select_dev(self.scores, self.offsets, self.choice, self.cap)
        e.step_dev(self.player, placement=self.choice, auto_reset=True, reward=self.reward, done=self.done, field=self.field,
                   vec=self.vec, piece_idx=self.piece, perspective=self.perspective)
        self.player.bitwise_xor_(1)       # strict alternation of sides (worker.py:96)

    def step(self):
        """one ply for every game; results in .reward / .done / .field / .vec / .piece (state after the ply, fused reset applied),
        the world model of the ply in .p_* / .K / .offsets / .scores / .choice.  Asynchronous."""
        torch = self.torch
        if not self._use_graph:
            self._ply()
        elif self._graph is None:
            if self.plies == 0:           # warm-up ply on a side stream: the library's lazy allocations happen here, not in the capture
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    self._ply()
                torch.cuda.current_stream().wait_stream(s)
            else:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._ply()
                self._graph = g
                g.replay()                # capture does not execute
        else:
            self._graph.replay()
        self.plies += 1

    def overflowed(self):
        """True if the last ply had more placements than the buffers hold (synchronises)."""
        return int(self.offsets[-1].item()) > self.cap
