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
  * post-processing (postproc_oracle.py) -- pinned against runs of the reference's own
    DBPostProcessor.__call__ (src/det/test.py imported unchanged by tests/golden/make_postproc_reference.py;
    the absent pyclipper / shapely replaced by clipper_offset.py, a restatement of Clipper 6.4.2's
    ClipperOffset without its final union): exact rectangles and scores on rectangular components, an
    agreement rate on noisy ones.  PARTIALLY PINNED: Clipper's union step / polygon vertex lists are not
    reproduced (the consumer only takes cv2.boundingRect of a box).
"""
