"""What bounds e2e: the PCIe link of the box.  Times the query upload (99.6 MB H2D from pinned memory) and the
response download (49.3 MB D2H) of the 2^24-vs-2^10 workload alone and concurrently, with CUDA events."""
import json
import torch

MB_H2D, MB_D2H = 99.6, 49.3


def timed(fn, streams, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    for _ in range(reps):
        fn()
    for s in streams:
        ev = torch.cuda.Event()
        ev.record(s)
        torch.cuda.current_stream().wait_event(ev)
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n_in, n_out = int(MB_H2D * 1e6 / 8), int(MB_D2H * 1e6 / 8)
    h_in = torch.empty(n_in, dtype=torch.int64, pin_memory=True)
    h_out = torch.empty(n_out, dtype=torch.int64, pin_memory=True)
    d_in = torch.empty(n_in, dtype=torch.int64, device="cuda")
    d_out = torch.empty(n_out, dtype=torch.int64, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d()
        d2h()
    t_in, t_out, t_both = timed(h2d, [s1]), timed(d2h, [s2]), timed(both, [s1, s2])
    print(json.dumps({"h2d_ms": t_in, "h2d_gbs": MB_H2D / t_in, "d2h_ms": t_out, "d2h_gbs": MB_D2H / t_out,
                      "concurrent_ms": t_both, "concurrent_h2d_gbs_floor": MB_H2D / t_both,
                      "note": "e2e per query cannot be below concurrent_ms: every query moves 99.6 MB up and 49.3 MB down"}))


if __name__ == "__main__":
    main()
