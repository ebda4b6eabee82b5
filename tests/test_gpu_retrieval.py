"""IndexFlatL2 (tcgen05 GEMM + fused top-k + fp32 re-rank) vs the fp32 oracle on planted, tie-free data."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_db,nq,d,k", [
    (5000, 300, 256, 10),      # ragged: 5000 % 256 != 0, 300 % 128 != 0
    (20000, 128, 4096, 25),    # V2-S VLAD width, BASELINE k
    (9000, 77, 1536, 20),      # V2-N width, faiss call-site k
    (3000, 10, 100, 5),        # d not a multiple of 64 (padded operand)
    (70000, 256, 512, 31),     # several strips, k at the limit
])
def test_flat_l2_matches_oracle(n_db, nq, d, k):
    from nano_vs_slam_b200.retrieval import IndexFlatL2
    from nano_vs_slam_b200.synthetic import planted_retrieval_set
    from oracle import glue_ref

    db, q, planted = planted_retrieval_set(n_db, nq, d, k, seed=n_db, device="cuda")
    index = IndexFlatL2(d)
    index.add(db)
    D, I = index.search(q, k)
    torch.cuda.synchronize()
    Dr, Ir = glue_ref.flat_l2_search(db.cpu(), q.cpu(), k)
    assert I.dtype == torch.int64
    assert torch.equal(I.cpu(), Ir), "top-k index lists must be bit-exact on tie-free data"
    assert torch.equal(I.cpu(), planted.cpu())
    np.testing.assert_allclose(D.cpu().numpy(), Dr.numpy(), rtol=1e-4, atol=2e-6)
    assert bool((D[:, 1:] >= D[:, :-1]).all())


def test_flat_l2_numpy_api_and_sharded_merge():
    from nano_vs_slam_b200.retrieval import IndexFlatL2, merge_topk_device, shard_bounds
    from nano_vs_slam_b200.synthetic import planted_retrieval_set
    from oracle import glue_ref

    n_db, nq, d, k = 12000, 200, 256, 20
    db, q, planted = planted_retrieval_set(n_db, nq, d, k, seed=1, device="cuda")
    index = IndexFlatL2(d)
    index.add(db.cpu().numpy())
    D, I = index.search(q.cpu().numpy(), k)          # faiss-shaped numpy call
    assert isinstance(D, np.ndarray) and I.dtype == np.int64 and np.array_equal(I, planted.cpu().numpy())
    # 8 shards searched in-process, merged like the all_gather result
    Ds, Is = [], []
    for r in range(8):
        lo, hi = shard_bounds(n_db, 8, r)
        sh = IndexFlatL2(d)
        sh.add(db[lo:hi])
        a, b = sh.search_device(q, k, id_offset=lo)
        Ds.append(a); Is.append(b)
    Dm, Im = merge_topk_device(torch.stack(Ds), torch.stack(Is))
    assert torch.equal(Im, planted)
    Dr, Ir = glue_ref.flat_l2_search(db.cpu(), q.cpu(), k)
    np.testing.assert_allclose(Dm.cpu().numpy(), Dr.numpy(), rtol=1e-4, atol=2e-6)
