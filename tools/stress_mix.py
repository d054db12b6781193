"""Development aid: does any kernel of the family leave state behind (shared memory, tensor memory) that changes the result
of a LATER launch of another kernel?  Builds one op of each kind at production shapes, records each op's output checksum
when run alone, then runs them in random order thousands of times and compares every output after every launch."""
import os
import random
import sys

sys.path.insert(0, os.environ.get("B200_PKG_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ocr_vi_invoice_b200 import ops


def checksum(t):
    v = t.contiguous().view(torch.int16) if t.element_size() == 2 else t.contiguous().view(torch.int32)
    return int(v.to(torch.int64).sum().item())


def conv(name, b, h, w, ci, co, k, s, flags=0, res=False, pair_ok=True):
    dev = "cuda"
    x = torch.randn(b, h, w, ci, device=dev).to(torch.bfloat16)
    n_pad = ops.choose_n_pad(co)
    wt = (torch.randn(n_pad, k * k * ci, device=dev) / (k * k * ci) ** 0.5).to(torch.bfloat16)
    sc, sh = torch.rand(n_pad, device=dev) + 0.5, torch.randn(n_pad, device=dev) * 0.1
    y = torch.empty(b, h // s, w // s, co, device=dev, dtype=torch.bfloat16)
    r = torch.randn(b, h // s, w // s, co, device=dev).to(torch.bfloat16) if res else None
    op = ops.Igemm(ops.IGEMM_CONV, x, wt, sc, sh, y, batch=b, in_h=h, in_w=w, in_c=ci, out_c=co, ksize=k, stride=s,
                   flags=flags | (ops.EPI_RES if res else 0), residual=r, res_ld=co if res else 0, name=name)
    return name, op, y


def chained(name, b, hw, n2):
    dev = "cuda"
    x = torch.randn(b, hw, hw, 64, device=dev).to(torch.bfloat16)
    wt = (torch.randn(256, 64, device=dev) / 8).to(torch.bfloat16)
    w2 = (torch.randn(n2, 256, device=dev) / 16).to(torch.bfloat16)
    sc, sh = torch.rand(256, device=dev) + 0.5, torch.randn(256, device=dev) * 0.1
    s2, t2 = torch.rand(n2, device=dev) + 0.5, torch.randn(n2, device=dev) * 0.1
    r = torch.randn(b, hw, hw, 256, device=dev).to(torch.bfloat16)
    y = torch.empty(b, hw, hw, 256, device=dev, dtype=torch.bfloat16)
    y2 = torch.empty(b, hw, hw, n2, device=dev, dtype=torch.bfloat16)
    op = ops.Igemm(ops.IGEMM_CONV, x, wt, sc, sh, y, batch=b, in_h=hw, in_w=hw, in_c=64, out_c=256, ksize=1, stride=1,
                   flags=ops.EPI_RELU | ops.EPI_RES, residual=r, res_ld=256, chain=(w2, s2, t2, y2, True), name=name)
    return name, op, y2


def projected(name, b, hw, c1, c2, co, s2, n2=0):
    """conv3 with the projection shortcut as second K segment (x2), optionally with the next conv1 chained on (no residual)."""
    dev = "cuda"
    x = torch.randn(b, hw, hw, c1, device=dev).to(torch.bfloat16)
    x2 = torch.randn(b, hw * s2, hw * s2, c2, device=dev).to(torch.bfloat16)
    wt = (torch.randn(co, c1 + c2, device=dev) / (c1 + c2) ** 0.5).to(torch.bfloat16)
    sc, sh = torch.ones(co, device=dev), torch.randn(co, device=dev) * 0.1
    y = torch.empty(b, hw, hw, co, device=dev, dtype=torch.bfloat16)
    chain, out = None, y
    if n2:
        wc = (torch.randn(n2, co, device=dev) / co ** 0.5).to(torch.bfloat16)
        y2 = torch.empty(b, hw, hw, n2, device=dev, dtype=torch.bfloat16)
        chain, out = (wc, torch.rand(n2, device=dev) + 0.5, torch.randn(n2, device=dev) * 0.1, y2, True), y2
    op = ops.Igemm(ops.IGEMM_CONV, x, wt, sc, sh, y, batch=b, in_h=hw, in_w=hw, in_c=c1, out_c=co, ksize=1, stride=1,
                   flags=ops.EPI_RELU, x2=x2, in_c2=c2, stride2=s2, chain=chain, name=name)
    return name, op, out


def dcn(name, b, hw, c, s):
    dev = "cuda"
    x = torch.randn(b, hw, hw, c, device=dev).to(torch.bfloat16)
    wt = torch.randn(c, c, 3, 3, device=dev) / (9 * c) ** 0.5
    oh = hw // s
    om = torch.zeros(b, oh, oh, 32, device=dev)
    om[..., :18] = torch.randn(b, oh, oh, 18, device=dev) * 2.0
    om[..., 18:27] = torch.randn(b, oh, oh, 9, device=dev)
    coords = torch.empty(b, oh, oh, 9, 4, device=dev, dtype=torch.int32)
    ops.dcn_coords(om, coords, b, hw, hw, c, s)
    sc, sh = torch.ones(c, device=dev), torch.zeros(c, device=dev)
    y = torch.empty(b, oh, oh, c, device=dev, dtype=torch.bfloat16)
    op = ops.Igemm(ops.IGEMM_DCN_TMEM, x, ops.pack_dcn_weight(wt, c), sc, sh, y, batch=b, in_h=hw, in_w=hw, in_c=c, out_c=c, ksize=3,
                   stride=s, flags=ops.EPI_RELU, offsets=coords, name=name)
    return name, op, y


def dcn_v1(name, b, hw, c, s):
    dev = "cuda"
    x = torch.randn(b, hw, hw, c, device=dev).to(torch.bfloat16)
    wt = torch.randn(c, c, 3, 3, device=dev) / (9 * c) ** 0.5
    oh = hw // s
    om = torch.zeros(b, oh, oh, 32, device=dev)
    om[..., :18] = torch.randn(b, oh, oh, 18, device=dev) * 2.0
    om[..., 18:27] = torch.randn(b, oh, oh, 9, device=dev)
    coords = torch.empty(b, oh, oh, 9, 4, device=dev, dtype=torch.int32)
    ops.dcn_coords(om, coords, b, hw, hw, c, s)
    n_pad = ops.choose_n_pad(c)
    sc, sh = torch.ones(n_pad, device=dev), torch.zeros(n_pad, device=dev)
    y = torch.empty(b, oh, oh, c, device=dev, dtype=torch.bfloat16)
    op = ops.Igemm(ops.IGEMM_DCN, x, ops.pack_conv_weight(wt, n_pad), sc, sh, y, batch=b, in_h=hw, in_w=hw, in_c=c, out_c=c, ksize=3,
                   stride=s, flags=ops.EPI_RELU, offsets=coords, name=name)
    return name, op, y


def offset_conv(name, b, hw, c, s):
    dev = "cuda"
    x = torch.randn(b, hw, hw, c, device=dev).to(torch.bfloat16)
    wt = (torch.randn(32, 9 * c, device=dev) / (9 * c) ** 0.5).to(torch.bfloat16)
    sc, sh = torch.ones(32, device=dev), torch.zeros(32, device=dev)
    y = torch.empty(b, hw // s, hw // s, 9, 4, device=dev, dtype=torch.int32)
    op = ops.Igemm(ops.IGEMM_CONV, x, wt, sc, sh, y, batch=b, in_h=hw, in_w=hw, in_c=c, out_c=32, ksize=3, stride=s,
                   flags=ops.EPI_DCN_COORDS, out_ld=32, name=name)
    return name, op, y


def head_tail(name, b, hw):
    dev = "cuda"
    t01 = torch.randn(b, hw, hw, 128, device=dev).to(torch.bfloat16)
    wt = (torch.randn(256, 64, device=dev) / 8).to(torch.bfloat16)
    sc, sh = torch.ones(256, device=dev), torch.randn(256, device=dev) * 0.1
    aux = torch.randn(257, device=dev)
    y = torch.empty(b, 1, 4 * hw, 4 * hw, device=dev)
    op = ops.Igemm(ops.IGEMM_CONV, t01[..., 64:], wt, sc, sh, y, batch=b, in_h=hw, in_w=hw, in_c=64, out_c=256, ksize=1, stride=1,
                   flags=ops.EPI_RELU | ops.EPI_SHUFFLE2 | ops.EPI_HEAD_DOT, shuf_c=64, x_ld=128, offsets=aux, name=name)
    op._extra = (t01, aux)
    return name, op, y


def asf(name, b, h2):
    dev = "cuda"
    sizes = [(h2, h2), (h2 // 2, h2 // 2), (h2 // 4, h2 // 4), (h2 // 8, h2 // 8)]
    ps = [torch.randn(b, h, w, 256, device=dev).to(torch.bfloat16) for h, w in sizes]
    zs = [torch.randn(b, h, w, 8, device=dev) for h, w in sizes]
    b_att = torch.randn(4, device=dev) * 0.1
    out = torch.empty(b, h2, h2, 256, device=dev, dtype=torch.bfloat16)
    return name, ops.AsfPlan(ps, zs, 8, b_att, out, b, sizes), out


def main():
    torch.manual_seed(0)
    random.seed(0)
    b = 16
    def maybe(fn, *a_):
        try:
            return [fn(*a_)]
        except (AttributeError, TypeError):  # an older build of the package without this operator
            return []
    items = [
        conv("layer1.conv3 64->256 res", b, 256, 256, 64, 256, 1, 1, ops.EPI_RELU, res=True),
        conv("layer1.conv1 256->64", b, 256, 256, 256, 64, 1, 1, ops.EPI_RELU),
        conv("layer1.conv2 3x3 64->64", b, 256, 256, 64, 64, 3, 1, ops.EPI_RELU),
        conv("layer1.down 64->256", b, 256, 256, 64, 256, 1, 1, 0),
        conv("smooth1 3x3 256->256 (pair)", b, 128, 128, 256, 256, 3, 1, ops.EPI_RELU),
        conv("layer3.conv1 1024->256 (pair)", b, 64, 64, 1024, 256, 1, 1, ops.EPI_RELU),
        conv("head.0 3x3 256->128 (pair + share)", b, 128, 128, 256, 128, 3, 1, ops.EPI_RELU),
        *maybe(dcn, "dcn c3", b, 128, 128, 1),
        *maybe(dcn, "dcn c4", b, 64, 256, 1),
        dcn_v1("dcn v1 c5 (N=512)", b, 32, 512, 1),
        dcn_v1("dcn v1 c3", b, 128, 128, 1),
        offset_conv("offset conv c3", b, 128, 128, 1),
        head_tail("head tail", b, 256),
        *maybe(asf, "asf tensor-core", b, 256),
        conv("layer2.conv3 128->512 res", b, 128, 128, 128, 512, 1, 1, ops.EPI_RELU, res=True),
        *maybe(chained, "chained conv3 -> conv1 (64)", b, 256, 64),
        *maybe(chained, "chained conv3 -> conv1 (128)", b, 256, 128),
        conv("layer2.conv1 512->128", b, 128, 128, 512, 128, 1, 1, ops.EPI_RELU),
        *maybe(projected, "projection in conv3, layer1.0 (chained)", b, 256, 64, 64, 256, 1, 64),
        *maybe(projected, "projection in conv3, layer3.0 (pair)", b, 64, 256, 512, 1024, 2),
    ]
    torch.cuda.synchronize()
    ref = {}
    for name, op, y in items:
        op()
        torch.cuda.synchronize()
        ref[name] = checksum(y)
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    only = sys.argv[2].split(",") if len(sys.argv) > 2 else None
    if only:
        items = [it_ for it_ in items if any(o in it_[0] for o in only)]
    refs = {name: y.clone() for name, op, y in items}
    bad = {}
    for it in range(iters):
        order = items[:]
        random.shuffle(order)
        for name, op, y in order:  # launched back to back, checked after the burst
            y.zero_()
            op()
        torch.cuda.synchronize()
        for k_, (name, op, y) in enumerate(order):
            if checksum(y) != ref[name]:
                bad[name] = bad.get(name, 0) + 1
                if bad[name] <= 6:
                    d = (y != refs[name])
                    ii = d.nonzero()
                    lo, hi = ii.min(0).values.tolist(), ii.max(0).values.tolist()
                    mag = (y.float() - refs[name].float()).abs()
                    print(f"iter {it}: {name} differs after '{order[k_ - 1][0] if k_ else '-'}': {int(d.sum())} elements, index range {lo}..{hi}, "
                          f"max |d| {float(mag.max()):.4g}", flush=True)
                    if "res" in name and bad[name] <= 3:
                        # which (tile, 64-channel chunk) regions, and does the wrong value look like "another residual"?
                        r = op._keep[5].float()
                        yw, yr = y.float(), refs[name].float()
                        B_, H_, W_, C_ = y.shape
                        tw = 128 if W_ % 128 == 0 else W_
                        dd = d.view(B_, H_, W_ // tw, tw, C_ // 64, 64)
                        cnt = dd.sum(dim=(3, 5))  # [B, H, W/tw, C/64]
                        for idx in cnt.nonzero().tolist()[:24]:
                            bi, hi_, wi, ci = idx
                            reg = (slice(bi, bi + 1), slice(hi_, hi_ + 1), slice(wi * tw, wi * tw + tw), slice(ci * 64, ci * 64 + 64))
                            w_, f_, rr = yw[reg], yr[reg], r[reg]
                            pos = (w_ > 0) & (f_ > 0)
                            delta = (w_ - f_)[pos]
                            # residual actually used = r + delta; find which other region's residual it equals
                            used = (rr[pos] + delta)
                            hits = []
                            for cj in range(C_ // 64):
                                for dh in (-74 * 2, -74, 0, 74, 74 * 2):
                                    h2 = hi_ + dh
                                    b2 = bi + (h2 // H_ if h2 >= 0 else -1)
                                    h2 = h2 % H_
                                    if 0 <= b2 < B_:
                                        cand = r[b2:b2 + 1, h2:h2 + 1, wi * tw:wi * tw + tw, cj * 64:cj * 64 + 64][pos]
                                        if cand.numel() and (cand - used).abs().max().item() < 0.07:
                                            hits.append((b2, h2, cj))
                            print(f"    tile (img {bi}, row {hi_}, col {wi * tw}) chunk {ci}: {int(cnt[bi, hi_, wi, ci])} of {tw * 64} differ; "
                                  f"residual used matches region(s) (img, row, chunk) {hits}; zero residual: {bool(used.abs().max().item() < 0.07) if used.numel() else None}",
                                  flush=True)
    print("no cross-kernel interference" if not bad else f"INTERFERENCE: {bad}")


if __name__ == "__main__":
    main()
