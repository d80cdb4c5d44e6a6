"""CUDA-graph capture of a fixed-shape step built from this package's ops.

Every launcher of libacfm_b200.so is capture-safe (no allocation, no synchronisation, shared-memory attributes set once
per device), and the autograd wrappers allocate through torch's caching allocator, which serves a private pool during
capture — so a whole forward + backward (+ optimiser) step can be recorded once and replayed with one launch.  This is what
removes the ~80 host-side launches of a training step from the critical path when the caller synchronises every step
(reads the loss), and what `predictor.PostOptimizer` does for its inner iteration.
"""
import torch


class CapturedStep:
    """Record `fn(*static_inputs)` once, replay it for new inputs of the same shapes.

        step = CapturedStep(fn, example_inputs)         # 3 eager warm-up calls, then capture
        outs = step(*inputs)                            # copies `inputs` (device or pinned host tensors) into the static
                                                        # buffers on the current stream, replays, returns the static outputs

    `fn` must be shape-static, free of host synchronisation and must not depend on Python-side state that changes between
    calls.  Outputs are the tensors (or tuple of tensors) `fn` returned during capture; they are overwritten by every
    replay — copy them out before the next call if they must survive."""

    def __init__(self, fn, example_inputs, warmup=3):
        self.fn = fn
        self.static_in = [t.detach().clone() if t.is_cuda else t.detach().to("cuda", copy=True) for t in example_inputs]
        dev = self.static_in[0].device
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(s)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs):
        for s, t in zip(self.static_in, inputs):
            if t is not s:
                s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.static_out


class PrefetchedStep:
    """Input pipeline around a CapturedStep: the pinned-host inputs of step i+1 are copied to the device on a copy stream
    while step i replays, the way a training loop prefetches its next batch (DataLoader(pin_memory=True) + non_blocking
    copies).  Two staging sets alternate; the replay takes its inputs from the staged set with device-to-device copies.

        pipe = PrefetchedStep(captured)
        pipe.prefetch(*host_inputs)                 # inputs of the first step
        for batch in batches:
            outs = pipe.run()                       # waits for the staged inputs only, then replays
            pipe.prefetch(*next_host_inputs)        # H2D of the next step overlaps this step's kernels
            ... read outs (D2H + synchronize) ...
    """

    def __init__(self, captured):
        self.captured = captured
        dev = captured.static_in[0].device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.staging = [[torch.empty_like(t) for t in captured.static_in] for _ in range(2)]
        self.ready = [torch.cuda.Event() for _ in range(2)]      # H2D into staging[i] finished
        self.consumed = [torch.cuda.Event() for _ in range(2)]   # staging[i] has been copied into the graph's inputs
        self.head = 0      # set the next prefetch writes
        self.pending = []  # staged sets not yet run, oldest first

    def prefetch(self, *host_inputs):
        if len(self.pending) == 2:
            raise RuntimeError("PrefetchedStep: both staging sets are full; call run() first")
        i = self.head
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[i])       # no-op for a never-recorded event
            for s, t in zip(self.staging[i], host_inputs):
                s.copy_(t, non_blocking=True)
            self.ready[i].record(self.copy_stream)
        self.pending.append(i)
        self.head = 1 - i

    def run(self):
        if not self.pending:
            raise RuntimeError("PrefetchedStep: run() without a prefetch()")
        i = self.pending.pop(0)
        cur = torch.cuda.current_stream(self.captured.static_in[0].device)
        cur.wait_event(self.ready[i])
        for s, t in zip(self.captured.static_in, self.staging[i]):
            s.copy_(t, non_blocking=True)
        self.consumed[i].record(cur)
        self.captured.graph.replay()
        return self.captured.static_out
