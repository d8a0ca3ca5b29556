"""GPU half of tests/test_datasets.py: ``Slice3DDataset.batch(indices, "cuda:0")`` -- the PNGs decoded on the host, the
image half of all 13 x B images in the library's two preprocessing kernels -- against the golden feed_dicts of the
reference's own class.  Bar: equal (8-bit image arithmetic; the camera / query tensors are host arithmetic copied over).

Written after the round's GPU budget was spent: the first run of this file is the round-end driver's (it sorts last so
that the parity tests proper run before it)."""
import numpy as np
import pytest
import torch

pytest.importorskip("PIL.Image")

from slice3d_b200.datasets import Slice3DDataset  # noqa: E402
from tests import dataset_files, helpers  # noqa: E402
from tests.test_datasets import KEYS, _golden_item  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("tag", list(dataset_files.CONFIGS))
def test_batch_on_device_equals_reference_items(tmp_path, tag):
    root = dataset_files.write(str(tmp_path))
    g = helpers.load_case("dataset_items")
    ds = Slice3DDataset("test", dataset_files.args(root, **dataset_files.CONFIGS[tag]))
    got = ds.batch([0, 1], "cuda:0")
    assert set(got) == KEYS and all(v.is_cuda for v in got.values())
    for i in range(2):
        want = _golden_item(g, tag, i)
        for k in KEYS:
            assert np.array_equal(got[k][i].cpu().numpy(), want[k]), (tag, i, k)
