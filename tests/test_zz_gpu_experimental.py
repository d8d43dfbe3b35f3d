"""Experimental kernels kept in the tree behind environment switches (run last: the persistent whole-RDB kernel uses
grid-wide barriers and needs all its CTAs co-resident — an idle GPU)."""
import pytest
import torch

from oracle import srn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('shape,chunk', [((3, 3, 32, 32), 2), ((2, 3, 40, 24), 1), ((5, 3, 48, 64), 3)])
def test_persistent_rdb_kernel_is_bit_identical_to_stage_launches(shape, chunk, monkeypatch):
    """dasr_rdb_tc (five dense-block stages inside one persistent kernel with grid barriers, experimental) must
    reproduce the five-launch schedule bit for bit: same MMAs in the same order, same epilogue arithmetic."""
    from dasr_b200 import engine
    nb = 2
    sd = O.synth_state_dict(O.rrdbnet_shapes(nb=nb), 1, 0.1)
    params = [v.cuda() for v in sd.values()]
    x = O.synth_image(shape, 2).cuda()
    monkeypatch.setenv('DASR_B200_RDB', '0')
    monkeypatch.setenv('DASR_B200_SCHED', '1')          # the persistent kernel implements the column-strip schedule
    ref = engine.rrdb_forward_bf16(x, params, nb, 4, engine._PackCache())
    monkeypatch.setenv('DASR_B200_RDB', '1')
    monkeypatch.setattr(engine, 'RDB_CHUNK_IMGS', chunk)
    got = engine.rrdb_forward_bf16(x, params, nb, 4, engine._PackCache())
    torch.cuda.synchronize()
    assert torch.equal(ref, got)
