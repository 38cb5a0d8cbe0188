"""CUDA-graph capture of a static-shape piece of the path (a decoder tower, a whole decode -> shade -> render step).

The layers between the reference's Python and the kernels issue tens of small launches per frame; at B = 1 the
low-resolution decoder layers and the per-Gaussian kernels are launch-bound, so replaying one captured graph per frame
is worth more than any single kernel optimisation there.  Requirements on `fn`: static shapes, no host
synchronisation (use the sync-free render path, `capacity=`), inputs and outputs are fixed tensors that the caller
refills / reads in place."""
import torch


class Graphed:
    def __init__(self, fn, warmup: int = 3, device=None):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        with torch.cuda.device(dev):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):  # allocator warm-up, cudaFuncSetAttribute calls, weight caches
                    fn()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = fn()

    def __call__(self):
        self.graph.replay()
        return self.out
