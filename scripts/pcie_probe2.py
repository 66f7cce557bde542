import time, torch
n = 16 << 20  # 16 MiB chunks
B = 8
h_in = torch.empty(B * n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(B * n, dtype=torch.uint8, pin_memory=True)
h_out2 = torch.empty(2 * B * n, dtype=torch.uint8, pin_memory=True)
d_in = [torch.empty(n, dtype=torch.uint8, device="cuda") for _ in range(3)]
d_out = [torch.empty(2 * n, dtype=torch.uint8, device="cuda") for _ in range(3)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(do_in, do_out, out_mult):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(B):
        if do_in:
            with torch.cuda.stream(s1):
                d_in[i % 3].copy_(h_in[i * n:(i + 1) * n], non_blocking=True)
        if do_out:
            with torch.cuda.stream(s2):
                m = out_mult * n
                (h_out2 if out_mult == 2 else h_out)[i * m:(i + 1) * m].copy_(d_out[i % 3][:m], non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3
for args in ((1, 0, 1), (0, 1, 1), (0, 1, 2), (1, 1, 1), (1, 1, 2)):
    run(*args)
    print(args, "ms", min(run(*args) for _ in range(5)))
