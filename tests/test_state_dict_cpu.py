"""CPU: checkpoint-loading behaviour of the conditional norm classes (ref pyfiles/model.py:24-52, 152-165).  Module
construction and load_state_dict need no kernel."""
import pytest
import torch

import cases


def test_cbinorm_drops_stale_running_stats_of_unversioned_checkpoints():
    model, _, _ = cases.use_product_modules()
    n = model.CBINorm2d(8, num_con=12, affine=True)
    sd = dict(n.state_dict())
    assert sorted(sd) == ["ConBias.0.bias", "ConBias.0.weight", "bias", "weight"]
    stale = dict(sd, running_mean=torch.zeros(8), running_var=torch.ones(8))
    with pytest.raises(RuntimeError, match="Unexpected running stats"):
        n.load_state_dict(stale)                         # no _metadata -> version None -> reported like the reference
    n.load_state_dict(sd)                                # a clean dict loads
    versioned = n.state_dict()                           # carries version-2 metadata
    assert versioned._metadata[""]["version"] == 2


def test_cbbnorm_defaults_num_batches_tracked_for_old_checkpoints():
    model, _, _ = cases.use_product_modules()
    n = model.CBBNorm2d(8, 12)
    assert sorted(n.state_dict()) == ["ConBias.0.bias", "ConBias.0.weight", "bias", "num_batches_tracked",
                                      "running_mean", "running_var", "weight"]
    n.num_batches_tracked.fill_(7)
    old = {k: v.clone() for k, v in n.state_dict().items() if k != "num_batches_tracked"}   # plain dict: no metadata
    n.load_state_dict(old)
    assert int(n.num_batches_tracked) == 0
    sd = n.state_dict()
    del sd["num_batches_tracked"]                        # version-2 metadata present: the key is required
    with pytest.raises(RuntimeError, match="num_batches_tracked"):
        n.load_state_dict(sd)
