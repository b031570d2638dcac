"""Do kernels on two streams / two branches of a CUDA graph overlap on this box?  (diagnostic, not a bench)"""
import torch
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
print(torch.cuda.get_device_name(0), "driver", torch.version.cuda, "n_dev", torch.cuda.device_count())
cyc = 2_000_000
def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
def one(): torch.cuda._sleep(cyc)
s2 = torch.cuda.Stream()
def two_streams():
    main = torch.cuda.current_stream()
    s2.wait_stream(main)
    torch.cuda._sleep(cyc)
    with torch.cuda.stream(s2):
        torch.cuda._sleep(cyc)
    main.wait_stream(s2)
one(); two_streams()
print("one sleep kernel      %.3f ms" % timed(one))
print("two streams, eager    %.3f ms" % timed(two_streams))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    main = torch.cuda.current_stream()
    br = torch.cuda.Stream()
    br.wait_stream(main)
    torch.cuda._sleep(cyc)
    with torch.cuda.stream(br):
        torch.cuda._sleep(cyc)
    main.wait_stream(br)
g.replay()
print("two graph branches    %.3f ms" % timed(g.replay))
