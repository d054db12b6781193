"""Times single igemm launches at production shapes (development aid, not part of the product)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ocr_vi_invoice_b200 import ops

def run(name, b, h, w, ci, co, k, s, flags, res=False, shuf=0, iters=20, block_n=0):
    dev = "cuda"
    x = torch.randn(b, h, w, ci, device=dev).to(torch.bfloat16)
    n_pad = ops.choose_n_pad(co)
    wt = (torch.randn(n_pad, k * k * ci, device=dev) / (k * k * ci) ** 0.5).to(torch.bfloat16)
    sc, sh = torch.ones(n_pad, device=dev), torch.zeros(n_pad, device=dev)
    oh, ow = h // s, w // s
    if shuf:
        y = torch.empty(b, 2 * oh, 2 * ow, shuf, device=dev, dtype=torch.bfloat16)
    else:
        y = torch.empty(b, oh, ow, co, device=dev, dtype=torch.bfloat16)
    r = torch.randn(b, oh, ow, co, device=dev).to(torch.bfloat16) if res else None
    op = ops.Igemm(ops.IGEMM_CONV, x, wt, sc, sh, y, batch=b, in_h=h, in_w=w, in_c=ci, out_c=co, ksize=k, stride=s,
                   flags=flags | (ops.EPI_RES if res else 0), residual=r, res_ld=co if res else 0, shuf_c=shuf, block_n=block_n, name=name)
    for _ in range(3):
        op()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        op()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    byts = x.numel() * 2 + y.numel() * 2 + (r.numel() * 2 if res else 0)
    if int(os.environ.get("B200_DEBUG", "0")) & 16:
        import ctypes as C
        from ocr_vi_invoice_b200 import _lib
        buf = (C.c_ulonglong * 16)()
        _lib.load().b200_igemm_debug_counters(buf)   # clear
        op(); torch.cuda.synchronize()
        _lib.load().b200_igemm_debug_counters(buf)
        names = ["A.empty", "B.empty", "mma.tempty", "mma.a_full", "mma.b_full", "dma.out_ready", "dma.wait_read", "epi.tfull", "epi.slot", "total", "mma.issue"]
        print("   " + "  ".join(f"{n}={buf[i]}" for i, n in enumerate(names)))
    print(f"{name:28s} {ms:7.3f} ms  {byts / ms / 1e9:7.2f} TB/s  {op.flops / ms / 1e9:8.1f} TF/s  {op.info()}", flush=True)

def run_head_dot(iters=20):
    dev = "cuda"
    b, h, w = 16, 256, 256
    t01 = torch.randn(b, h, w, 128, device=dev).to(torch.bfloat16)
    wt = (torch.randn(256, 64, device=dev) / 8).to(torch.bfloat16)
    sc, sh = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    aux = torch.randn(257, device=dev)
    y = torch.empty(b, 1, 4 * h, 4 * w, device=dev)
    op = ops.Igemm(ops.IGEMM_CONV, t01[..., 64:], wt, sc, sh, y, batch=b, in_h=h, in_w=w, in_c=64, out_c=256, ksize=1, stride=1,
                   flags=ops.EPI_RELU | ops.EPI_SHUFFLE2 | ops.EPI_HEAD_DOT, shuf_c=64, x_ld=128, offsets=aux, name="head-dot")
    for _ in range(3):
        op()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        op()
    e1.record()
    torch.cuda.synchronize()
    print(f"head-dot convT   {e0.elapsed_time(e1) / iters:7.3f} ms  {op.info()}", flush=True)


def run_offset(name, b, h, w, ci, s, iters=20):
    """offset_mask_conv (3x3 -> 27, N padded to 32) with the sampling-record epilogue."""
    dev = "cuda"
    x = torch.randn(b, h, w, ci, device=dev).to(torch.bfloat16)
    wt = (torch.randn(32, 9 * ci, device=dev) / (9 * ci) ** 0.5).to(torch.bfloat16)
    sc, sh = torch.ones(32, device=dev), torch.zeros(32, device=dev)
    oh, ow = h // s, w // s
    y = torch.empty(b, oh, ow, 9, 4, device=dev, dtype=torch.int32)
    op = ops.Igemm(ops.IGEMM_CONV, x, wt, sc, sh, y, batch=b, in_h=h, in_w=w, in_c=ci, out_c=32, ksize=3, stride=s,
                   flags=ops.EPI_DCN_COORDS, out_ld=32, name=name)
    for _ in range(3):
        op()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        op()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if int(os.environ.get("B200_DEBUG", "0")) & 16:
        import ctypes as C
        from ocr_vi_invoice_b200 import _lib
        buf = (C.c_ulonglong * 16)()
        _lib.load().b200_igemm_debug_counters(buf)
        op(); torch.cuda.synchronize()
        _lib.load().b200_igemm_debug_counters(buf)
        names = ["A.empty", "B.empty", "mma.tempty", "mma.a_full", "mma.b_full", "dma.out_ready", "dma.wait_read", "epi.tfull", "epi.slot", "total", "mma.issue"]
        print("   " + "  ".join(f"{n}={buf[i]}" for i, n in enumerate(names)))
    print(f"{name:28s} {ms:7.3f} ms  {op.flops / ms / 1e9:8.1f} TF/s  {op.info()}", flush=True)


if __name__ == "__main__" and os.environ.get("EPI_SET") == "offset":
    run_offset("layer2 offset s1", 16, 128, 128, 128, 1)
    run_offset("layer2 offset s2", 16, 256, 256, 128, 2)
    run_offset("layer3 offset s1", 16, 64, 64, 256, 1)
    run_offset("layer4 offset s1", 16, 32, 32, 512, 1)
    sys.exit(0)
if __name__ == "__main__" and os.environ.get("EPI_SET") == "conv3":
    for bn in (256, 128):
        run(f"layer2.conv3 128->512 res bn{bn}", 16, 128, 128, 128, 512, 1, 1, ops.EPI_RELU, res=True, block_n=bn)
        run(f"layer3.conv3 256->1024 res bn{bn}", 16, 64, 64, 256, 1024, 1, 1, ops.EPI_RELU, res=True, block_n=bn)
        run(f"layer3.conv3 no-res bn{bn}", 16, 64, 64, 256, 1024, 1, 1, ops.EPI_RELU, block_n=bn)
        run(f"layer4.conv3 512->2048 res bn{bn}", 16, 32, 32, 512, 2048, 1, 1, ops.EPI_RELU, res=True, block_n=bn)
        run(f"layer4.conv3 no-res bn{bn}", 16, 32, 32, 512, 2048, 1, 1, ops.EPI_RELU, block_n=bn)
    sys.exit(0)
def run_stem(iters=20):
    dev = "cuda"
    b, h, w = 16, 1024, 1024
    img = torch.randn(b, 3, h, w, device=dev)
    wt = torch.randn(64, 3, 7, 7, device=dev) / 147 ** 0.5
    packed = torch.empty(b, h // 2 + 3, w // 2 + 3, 16, device=dev, dtype=torch.bfloat16)
    ops.stem_pack(img, packed)
    y = torch.empty(b, h // 2, w // 2, 64, device=dev, dtype=torch.bfloat16)
    op = ops.Igemm(ops.IGEMM_STEM, packed, ops.pack_stem_weight(wt, 64), torch.ones(64, device=dev), torch.zeros(64, device=dev), y,
                   batch=b, in_h=h, in_w=w, in_c=3, out_c=64, ksize=7, stride=2, flags=ops.EPI_RELU, name="stem")
    for _ in range(3):
        op()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        op()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    if int(os.environ.get("B200_DEBUG", "0")) & 16:
        import ctypes as C
        from ocr_vi_invoice_b200 import _lib
        buf = (C.c_ulonglong * 16)()
        _lib.load().b200_igemm_debug_counters(buf)
        op(); torch.cuda.synchronize()
        _lib.load().b200_igemm_debug_counters(buf)
        names = ["A.empty", "B.empty", "mma.tempty", "mma.a_full", "mma.b_full", "dma.out_ready", "dma.wait_read", "epi.tfull", "epi.slot", "total", "mma.issue"]
        print("   " + "  ".join(f"{n}={buf[i]}" for i, n in enumerate(names)))
    print(f"stem 7x7 s2 3->64 @1024      {ms:7.3f} ms  {op.flops / ms / 1e9:8.1f} TF/s  {op.info()}", flush=True)


if __name__ == "__main__" and os.environ.get("EPI_SET") == "stem":
    run_stem()
    sys.exit(0)
if __name__ == "__main__" and os.environ.get("EPI_SET") == "conv1":
    run("layer3.conv1 1024->256 64x64", 16, 64, 64, 1024, 256, 1, 1, ops.EPI_RELU)
    run("layer4.conv1 2048->512 32x32", 16, 32, 32, 2048, 512, 1, 1, ops.EPI_RELU)
    run("layer2.conv1 512->128 128x128", 16, 128, 128, 512, 128, 1, 1, ops.EPI_RELU)
    run("layer3.down 512->1024 s2", 16, 128, 128, 512, 1024, 1, 2, 0)
    run("lat1 512->256 128x128", 16, 128, 128, 512, 256, 1, 1, 0)
    sys.exit(0)
if __name__ == "__main__" and os.environ.get("EPI_SET") == "head":
    run_head_dot()
    sys.exit(0)
if __name__ == "__main__" and os.environ.get("EPI_SET") == "gemm":
    for n in (64, 128, 256):
        run(f"1x1 K=2048 N={n} 32x32", 16, 32, 32, 2048, n, 1, 1, 0, iters=50)
    for n in (64, 256):
        run(f"1x1 K=2048 N={n} 8x128", 16, 8, 128, 2048, n, 1, 1, 0, iters=50)
    for n in (64, 256):
        run(f"1x1 K=512 N={n} 64x64", 16, 64, 64, 512, n, 1, 1, 0, iters=50)
    sys.exit(0)
if __name__ == "__main__" and os.environ.get("EPI_SET") == "conv2":
    run("layer1.conv2 3x3 64->64", 16, 256, 256, 64, 64, 3, 1, ops.EPI_RELU)
    run("conv 3x3 64->64 no-relu", 16, 256, 256, 64, 64, 3, 1, 0)
    run("stem-like 1x1 k=4 256->64", 16, 512, 512, 256, 64, 1, 1, ops.EPI_RELU)
    run("3x3 64->128", 16, 256, 256, 64, 128, 3, 1, ops.EPI_RELU)
    run("3x3 64->256", 16, 256, 256, 64, 256, 3, 1, ops.EPI_RELU)
    run("3x3 128->64", 16, 256, 256, 128, 64, 3, 1, ops.EPI_RELU)
    sys.exit(0)
if __name__ == "__main__":
    run("layer1.conv3 64->256 res", 16, 256, 256, 64, 256, 1, 1, ops.EPI_RELU, res=True)
    run("layer1.conv3 no-res", 16, 256, 256, 64, 256, 1, 1, ops.EPI_RELU)
    run("convT 64->4x64", 16, 256, 256, 64, 256, 1, 1, ops.EPI_RELU | ops.EPI_SHUFFLE2, shuf=64)
    run("layer2.conv3 128->512 res", 16, 128, 128, 128, 512, 1, 1, ops.EPI_RELU, res=True)
    run("layer1.conv1 256->64", 16, 256, 256, 256, 64, 1, 1, ops.EPI_RELU)
    run("layer1.conv2 3x3 64->64", 16, 256, 256, 64, 64, 3, 1, ops.EPI_RELU)
