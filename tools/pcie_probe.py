#!/usr/bin/env python3
"""Host<->device copy ceiling for the e2e leg of bench.py: pinned 128 MiB buffers (one 2^24-word
transform), H2D alone, D2H alone, and both directions at once on two streams.  Prints one JSON line.
The e2e step moves 128 MiB each way, so `both_ms` is the floor of bench.py's e2e ms_per_step."""
import json

import torch


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    n = 1 << 24
    h_in = torch.empty(n, dtype=torch.int64).pin_memory()
    h_out = torch.empty(n, dtype=torch.int64).pin_memory()
    d_a = torch.empty(n, dtype=torch.int64, device="cuda")
    d_b = torch.empty(n, dtype=torch.int64, device="cuda")
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    nbytes = n * 8

    def h2d():
        d_a.copy_(h_in, non_blocking=True)

    def d2h():
        h_out.copy_(d_b, non_blocking=True)

    def both():
        cur = torch.cuda.current_stream()
        s_in.wait_stream(cur)
        s_out.wait_stream(cur)
        with torch.cuda.stream(s_in):
            d_a.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out.copy_(d_b, non_blocking=True)
        cur.wait_stream(s_in)
        cur.wait_stream(s_out)

    t_in, t_out, t_both = timed(h2d), timed(d2h), timed(both)
    print(json.dumps({"bytes_each_way": nbytes, "h2d_ms": t_in, "h2d_gbs": nbytes / t_in / 1e6,
                      "d2h_ms": t_out, "d2h_gbs": nbytes / t_out / 1e6, "both_ms": t_both,
                      "both_gbs_each_way": nbytes / t_both / 1e6}))


if __name__ == "__main__":
    main()
