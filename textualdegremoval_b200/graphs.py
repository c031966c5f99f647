"""CUDA-graph replay of an inference forward.

The forward of every network in ``archs`` is a fixed sequence of launches of libtdr_sm100.so on the current stream:
no host read-back, no data-dependent control flow (the MASA window origins and match indices stay on the device).
For a fixed input shape the whole sequence can therefore be captured once and replayed with one ``cudaGraphLaunch``:
at small batches (serving one image at a time) the coarse pyramid levels are launch-bound -- a guided Restormer
forward is ~570 launches -- and the replay removes the per-launch host cost and most of the inter-kernel gaps.

    fwd = GraphedForward(net, lq, ref)      # warm-up + capture on the shapes / dtypes of the example inputs
    y = fwd(lq2, ref2)                      # copies into the captured input buffers, replays, returns the output

The returned tensor is the graph's static output buffer: consume (or clone) it before the next call.  Results are
bit-identical to the eager path (same kernels, same launch parameters).  Inference only: the training step keeps
its explicit schedule (the gradient all-reduce is issued from the host as modules finish)."""
import torch


class GraphedForward:
    def __init__(self, net, *example_inputs, warmup: int = 2):
        assert all(isinstance(x, torch.Tensor) and x.is_cuda for x in example_inputs), "CUDA tensors only"
        self.net = net
        self.static_in = [x.clone() for x in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(1, warmup)):          # first-use work (function attributes, packed weights) happens here
                net(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = net(*self.static_in)

    def __call__(self, *inputs):
        assert len(inputs) == len(self.static_in)
        for dst, src in zip(self.static_in, inputs):
            if src.shape != dst.shape or src.dtype != dst.dtype:
                raise ValueError(f"GraphedForward was captured for {tuple(dst.shape)} {dst.dtype}, "
                                 f"got {tuple(src.shape)} {src.dtype}")
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.static_out
