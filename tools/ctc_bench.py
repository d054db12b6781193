import sys, torch, numpy as np
sys.path.insert(0, '.')
from ocr_vi_invoice_b200.model.rec2 import VOCAB, GreedyCTCDecoder, Tokenizer
dec = GreedyCTCDecoder(Tokenizer(VOCAB))
lp = torch.randn(64, 960, 232, device='cuda').log_softmax(-1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
for _ in range(3): dec.run_device(lp)
ms = []
for _ in range(20):
    flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dec.run_device(lp); b.record(); b.synchronize(); ms.append(a.elapsed_time(b))
print("ctc_greedy ms", float(np.median(ms)))
