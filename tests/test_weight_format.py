"""cb_load_weights takes the reference's flat weight export (python/export_weights_cuda.py:73-107).  Where the
reference checkout is present, its exporter's own build_parts() — imported in place — must produce, for the
reference model loaded with our seeded weights, exactly the blob oracle/weights.flatten() builds (the blob every
GPU test and the bench load), and cb_weight_count must agree with the exporter's EXPECTED_FLOATS."""
import os
import sys

import numpy as np
import pytest

from oracle import weights as W

REF_PY = "/root/reference/python"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF_PY), reason="reference checkout not present (GPU box)")


def test_blob_equals_the_reference_exporter(cb):
    from oracle.make_golden import reference_model
    sys.path.insert(0, REF_PY)
    try:
        import export_weights_cuda as ref_export
    finally:
        sys.path.remove(REF_PY)
    w = W.make_weights(6, 128, 1803, "B")
    model, _traced = reference_model(6, 128, w)
    ref_blob = np.concatenate(ref_export.build_parts(model)).astype(np.float32)
    ours = W.flatten(w, 6, 128)
    assert ref_blob.size == ref_export.EXPECTED_FLOATS == ours.size == W.blob_size(6, 128)
    assert np.array_equal(ref_blob, ours)
    L = cb.load_library()
    assert L.cb_weight_count(6, 128) == ref_export.EXPECTED_FLOATS
    assert W.unflatten(ours, 6, 128).keys() == w.keys()
