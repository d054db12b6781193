"""CPU restatement of the reference's DBNet++ detection path -- TEST INFRASTRUCTURE ONLY.

Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / `--impl reference` legs may import it, and only as the checker or the reported CPU
baseline.  The product path (ocr_vi_invoice_b200/) never imports this package and fails loudly when
its CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):
  * model forward (dbnet_oracle.py)  -- pinned against outputs of the reference's own modules
    (/root/reference/model/det, imported unchanged) by tests/golden/make_golden.py -> tests/golden/*.npz
  * deformable conv (deform_conv2d_ref) -- pinned against torchvision.ops.deform_conv2d (the third-party
    op the reference calls at model/det/dcn.py:48; unpinned version, 0.26.0 installed) via golden vectors
  * post-processing (postproc_oracle.py) -- PARITY UNPINNED against the reference's own boxes: the
    reference path needs pyclipper + shapely (src/det/test.py:17-18), neither is installed and there is
    no network.  It is pinned instead against the cv2 half of the reference path (findContours /
    approxPolyDP / box_score_fast run here, golden in tests/golden/postproc_cv2.npz) as an agreement
    rate, not bit-exactness.
"""
