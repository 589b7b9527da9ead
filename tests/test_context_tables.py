"""CPU suite: host logic of the context model -- inverse hash tables, stream chunking, python hash twin."""
import numpy as np
import torch

from conftest import R3


def test_python_hash_twin_matches_oracle(oracle, golden):
    from cnc_b200.context_models import get_grid_index

    rng = np.random.default_rng(0)
    for D, res, T in ((3, 18, 5832), (3, 148, 2 ** 19), (3, 514, 2 ** 19), (2, 1026, 2 ** 17), (2, 130, 16904), (3, 33, 35944)):
        pos = rng.integers(0, res, (5000, D))
        got = get_grid_index(T, res, torch.from_numpy(pos)).numpy()
        want = oracle.grid_rows(pos.astype(np.uint32), T, res)
        np.testing.assert_array_equal(got, want.astype(np.int64))


def test_inverse_tables_small_cpu():
    from cnc_b200.context_models import CNC_context_models, get_grid_index

    torch.manual_seed(0)
    res3, log2T = [6, 10, 18, 34], 12
    cm = CNC_context_models(num_dim=3, resolutions_list=res3, resolutions_list_2D=[18, 34], log2_hashmap_size=log2T,
                            log2_hashmap_size_2D=10, n_features=8, sample_num=500, ste_binary=True, Rb=16,
                            skip_levels_3D=(0, 1, 2), device="cpu")
    assert cm.offs == [0, 216, 1216, 5312, 9408]
    assert cm.resolution_thresh == 10.0 and cm.n_levels_thresh == 2
    for n, r in enumerate(res3):
        T = cm.offs[n + 1] - cm.offs[n]
        uv, cnt, cs, pts = cm.unique_value_list[n], cm.unique_count_list[n], cm.unique_count_cumsum_list[n], cm.pos_grid_sorted_list[n]
        assert int(cnt.sum()) == r ** 3 == pts.shape[0] and int(cs[-1]) == r ** 3 and cs[0] == 0
        assert uv.unique().numel() == uv.numel()
        # every voxel of a group hashes to the group's row
        rows = get_grid_index(T, r, pts.to(torch.int64))
        assert torch.equal(rows, torch.repeat_interleave(uv, cnt))
        if r ** 3 <= 2 ** log2T:
            assert (cnt == 1).all() and uv.numel() == r ** 3      # dense: shuffled symbol order, one voxel per row
            assert not torch.equal(uv, uv.sort().values)
        else:
            assert torch.equal(uv, uv.sort().values) and cnt.max() > 1  # hashed: ascending row order, collisions
    assert cm.ttl_hashparams_num_valid_levels == int(cm.hashparams_num_levels[3])
    assert len(cm._chunks(3)) == 1 and cm._chunks(3)[0] == (0, int(cm.hashparams_num_levels[3]))
    assert len(cm.context_model_2D) == 1 and cm.context_model_2D[0][0].in_features == 8 * 2 + 1
    assert cm.context_model_3D[0].in_features == 25


def test_stream_chunking_matches_reference_numbers():
    """at the product config the chunk size is min(floor(2e7 / (res^3 / entries)), entries) with the ratio
    evaluated in float32 like the reference (int64 / int64 tensor division) -> 504198 / 197258 / 77216 rows."""
    for res, entries, want in ((275, 524288, 504198), (376, 524288, 197258), (514, 524288, 77216), (201, 524288, 524288)):
        ratio = ((torch.tensor(res) ** 3) / torch.tensor(entries)).item()
        per = min(int(20000000 // ratio), entries)
        assert per == want
        assert -(-entries // per) == {504198: 2, 197258: 3, 77216: 7, 524288: 1}[want]
