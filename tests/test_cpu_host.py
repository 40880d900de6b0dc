"""Host-side logic of the Conos mirror (no GPU): argument checks with the reference's messages, gene-set
selection, pair partitioning, and the world_size-2 gloo allgatherv plumbing."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_argument_validation(small_panel):
    import conos_b200
    con = conos_b200.Conos(small_panel)
    with pytest.raises(ValueError, match="only the following spaces are currently supported"):
        con.buildGraph(space="UMAP")
    with pytest.raises(ValueError, match="matching methods"):
        con.buildGraph(matching_method="xNN")
    with pytest.raises(ValueError, match="distance metrics"):
        con.buildGraph(metric="cosine")
    with pytest.raises(ValueError, match="k1 must be >= k"):
        con.buildGraph(k=15, k1=10)
    with pytest.raises(NotImplementedError):
        con.buildGraph(space="JNMF")
    with pytest.raises(ValueError, match="named list"):
        conos_b200.Conos([1, 2])
    with pytest.raises(ValueError, match="identical to already existing"):
        con.addSamples({list(small_panel)[0]: list(small_panel.values())[0]})


def test_common_od_genes_matches_oracle(oracle, small_panel):
    import conos_b200
    from conftest import to_oracle_samples
    con = conos_b200.Conos(small_panel)
    S = to_oracle_samples(oracle, small_panel)
    names = list(small_panel)
    for n_od in (300, 50, 7):
        for a in range(3):
            for b in range(a + 1, 3):
                ids, ga, gb = con._common_od_genes(names[a], names[b], n_od)
                want = oracle.common_overdispersed_genes([S[names[a]], S[names[b]]], n_od)
                assert con._gene_names(ids) == want
                assert [small_panel[names[a]].genes[t] for t in ga] == want
                assert [small_panel[names[b]].genes[t] for t in gb] == want


def test_common_od_genes_partial_overlap(oracle):
    import scipy.sparse as sp
    import conos_b200
    rng = np.random.default_rng(0)

    def mk(genes, od):
        n = 6
        return conos_b200.Pagoda2(counts=sp.csc_matrix(rng.random((n, len(genes)))), genes=genes,
                                  cells=[f"c{t}" for t in range(n)], varinfo_v=np.zeros(len(genes)),
                                  varinfo_qv=np.ones(len(genes)), pca=rng.random((n, 3)), od_genes=od)
    a = mk(["B", "a", "C", "d", "E"], ["E", "a", "C", "B"])
    b = mk(["d", "C", "E", "Z", "a"], ["Z", "C", "d", "E"])
    con = conos_b200.Conos({"x": a, "y": b})
    ids, ga, gb = con._common_od_genes("x", "y", 3)
    S = {"x": oracle.Sample(a.counts, a.genes, a.cells, a.varinfo_v, a.varinfo_qv, a.pca, a.od_genes),
         "y": oracle.Sample(b.counts, b.genes, b.cells, b.varinfo_v, b.varinfo_qv, b.pca, b.od_genes)}
    want = oracle.common_overdispersed_genes([S["x"], S["y"]], 3)
    assert con._gene_names(ids) == want == ["C", "E", "a"]   # count 2 first (name order), 'Z'/'B' absent in one sample


def test_partition_pairs_lpt_with_locality():
    """cg_partition_pairs (GPU-free): balanced loads, every pair owned, fewer samples per rank than a contiguous cut."""
    from conos_b200.dist import partition_pairs
    for S, world in [(8, 8), (8, 2), (100, 8), (3, 8), (2, 1)]:
        pairs = [(i, j) for i in range(S) for j in range(i + 1, S)]
        a, b = [p[0] for p in pairs], [p[1] for p in pairs]
        cost = [1.0] * len(pairs)
        owner = partition_pairs(a, b, cost, world, sample_cost=[0.4] * S, n_samples=S)
        assert owner.min() >= 0 and owner.max() < world
        loads = np.bincount(owner, minlength=world)
        assert loads.max() - loads[loads > 0].min() <= max(2, 0.05 * loads.max())
        assert np.array_equal(owner, partition_pairs(a, b, cost, world, sample_cost=[0.4] * S, n_samples=S))
        if S == 100:   # locality: a rank sees far fewer than all 100 samples' worth of distinct partners
            touched = [len({a[q] for q in np.flatnonzero(owner == r)} | {b[q] for q in np.flatnonzero(owner == r)})
                       for r in range(world)]
            assert max(touched) <= 60, touched
    # unequal costs: the largest pair alone on its rank when it dominates
    owner = partition_pairs([0, 0, 1], [1, 2, 2], [100.0, 1.0, 1.0], 2)
    assert owner[0] != owner[1] and owner[1] == owner[2]


def _gloo_worker(rank, world, port, out_q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from conos_b200.dist import gathered_layout, partition_pairs, permute_gathered
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    S = 5
    pairs = [(i, j) for i in range(S) for j in range(i + 1, S)]
    owner = partition_pairs([p[0] for p in pairs], [p[1] for p in pairs], [float((i + 1) * (j + 1)) for i, j in pairs],
                            world, sample_cost=[1.0] * S, n_samples=S)
    mine = [q for q in range(len(pairs)) if owner[q] == rank]
    # every pair q contributes q + 1 rows (from = q, to = row)
    local = np.array([(q, r) for q in mine for r in range(q + 1)], np.int64).reshape(-1, 2)
    counts_local = np.zeros(len(pairs), np.int64)
    for q in mine:
        counts_local[q] = q + 1
    # the library does: allreduce(sum) of the counts, broadcast of every rank's table into the staging table
    t = torch.from_numpy(counts_local.copy())
    dist.all_reduce(t)
    counts = t.numpy()
    tables = [None] * world
    dist.all_gather_object(tables, local)
    rank_off, src_off, dst_off = gathered_layout(owner, counts, world)
    staging = np.concatenate(tables)
    assert [len(x) for x in tables] == np.diff(rank_off).tolist()
    out_q.put((rank, owner.tolist(), permute_gathered(staging, src_off, dst_off).tolist()))
    dist.destroy_process_group()


def test_allgatherv_layout_two_ranks_gloo():
    """The N > 1 host logic on CPU: two gloo ranks compute the same partition, and the gathered layout restores the
    reference's row order (pairs in combn order) whatever the partition was."""
    import torch.multiprocessing as mp
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctxmp.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    want = [[q_, r] for q_ in range(10) for r in range(q_ + 1)]
    assert res[0][1] == res[1][1] and len(set(res[0][1])) == 2       # same partition on both ranks, both ranks used
    for rank, owner, table in res:
        assert table == want


def test_lazy_containers():
    """PairEntry materialises the gene names on first access; LazyNames gathers the barcodes in vertex order and keeps
    its own copy of the vertex ids (the graph arrays are views into buffers reused by the next buildGraph)."""
    import numpy as np
    from conos_b200.conos import LazyNames, PairEntry
    calls = []

    def namer(ids):
        calls.append(1)
        return [f"g{t}" for t in ids]

    e = PairEntry(np.array([3, 1, 2]), namer, CPC=np.zeros((3, 2)))
    assert "genes" in e and "CPC" in e and not calls
    assert e["genes"] == ["g3", "g1", "g2"] and e["genes"] == ["g3", "g1", "g2"] and len(calls) == 1
    cells = np.array(["a", "b", "c", "d"], dtype=object)
    v = np.array([2, 0, 3], np.int32)
    ln = LazyNames(cells, v)
    v[:] = 0                      # the caller's buffer is overwritten by the next call
    assert len(ln) == 3 and ln[0] == "c" and list(ln) == ["c", "a", "d"]
    assert np.asarray(ln).tolist() == ["c", "a", "d"]


def test_gene_universe_mapping_host_side():
    """The upload universe is the union of the samples' top-n od lists; pair gene columns are remapped to the columns of
    the uploaded (subset) matrices.  Pure host logic: checked without a device by stubbing the upload."""
    import numpy as np
    import scipy.sparse as sp
    from conos_b200.conos import Conos, Pagoda2
    rng = np.random.default_rng(0)
    genes_a = [f"g{t:02d}" for t in range(10)]
    genes_b = [f"g{t:02d}" for t in (9, 8, 7, 6, 5, 4, 3, 2, 11, 12)]
    mk = lambda n, genes, od: Pagoda2(counts=sp.random(n, len(genes), 0.5, "csc", random_state=1), genes=genes,
                                      cells=[f"c{n}_{t}" for t in range(n)], varinfo_v=np.zeros(len(genes)),
                                      varinfo_qv=np.ones(len(genes)), pca=rng.normal(size=(n, 3)), od_genes=od)
    # build the vocabulary and membership without touching the device
    con = object.__new__(Conos)
    con.samples = {"A": mk(6, genes_a, ["g03", "g05", "g00", "g09"]), "B": mk(7, genes_b, ["g05", "g12", "g02", "g07"])}
    con._vocab, con._od_membership, con._panel_col = None, {}, {}
    ids, ga, gb = con._common_od_genes("A", "B", 3)
    names = con._gene_names(ids)
    # top-3 lists: A {g03,g05,g00}, B {g05,g12,g02}; table sorted by count then name; g12 is absent from A, g00 from B
    assert names == ["g05", "g02", "g03"]
    assert [genes_a[c] for c in ga] == names and [genes_b[c] for c in gb] == names
    # with a subset panel the columns are those of the uploaded matrices
    pc_a = np.full(10, -1, np.int32); pc_a[[0, 2, 3, 5]] = np.arange(4)       # uploaded: g00 g02 g03 g05
    pc_b = np.full(10, -1, np.int32); pc_b[[4, 6, 7, 9]] = np.arange(4)       # uploaded: g05 g03 g02 g12
    con._panel_col = {"A": pc_a, "B": pc_b}
    ids2, ga2, gb2 = con._common_od_genes("A", "B", 3)
    assert ids2.tolist() == ids.tolist()
    assert ga2.tolist() == [3, 1, 2] and gb2.tolist() == [0, 2, 1]


def test_od_genes_uniformly_host_side():
    """getOdGenesUniformly (R/conos.R:505-513): genes by their best rank in any sample's od list, ties in gene-name
    order (group_by sorts the groups, arrange(rank) is stable), first n.genes."""
    import numpy as np
    import scipy.sparse as sp
    from conos_b200.conos import Conos, Pagoda2
    mk = lambda od: Pagoda2(counts=sp.csc_matrix((2, 6)), genes=list("abcdef"), cells=["x", "y"],
                            varinfo_v=np.zeros(6), varinfo_qv=np.ones(6), pca=None, od_genes=od)
    con = object.__new__(Conos)
    con.samples = {"A": mk(["d", "a", "f"]), "B": mk(["b", "a", "c"]), "C": mk(["c", "e"])}
    # best ranks: d 0, b 0, c 0 (from C), a 1, e 1, f 2 -> ties by name
    assert con._od_genes_uniformly(10) == ["b", "c", "d", "a", "e", "f"]
    assert con._od_genes_uniformly(4) == ["b", "c", "d", "a"]


def test_propagate_labels_argument_checks_without_device():
    """Conos$propagateLabels: unknown method is an error with the reference's message (R/conclass.R:915-917); no graph
    is an error before any device call."""
    import pytest
    from conos_b200.conos import Conos
    con = object.__new__(Conos)
    con.graph = None
    with pytest.raises(ValueError, match="Only 'solver' and 'diffusion' are supported"):
        con.propagateLabels({}, method="magic")
    with pytest.raises(NotImplementedError):
        con.propagateLabels({}, method="solver")
    with pytest.raises(ValueError, match="buildGraph"):
        con.propagateLabels({})
    with pytest.raises(ValueError, match="buildGraph"):
        con.correctGenes()


def test_scanpy_graph_matrices_host_side():
    """ConosGraph.scanpy_matrices: as_adjacency_matrix(graph, attr='weight')[cell.ids, cell.ids] as dgCMatrix slots and
    distances 1 - x (R/integrations.R:170-180); the HDF5 writer asks for h5py like the reference asks for rhdf5."""
    import numpy as np
    import pytest
    from conos_b200.conos import ConosGraph
    names = np.array(["c2", "c0", "c3", "c1"], dtype=object)            # vertex order = first appearance
    ef, et = np.array([0, 0, 1], np.int32), np.array([1, 3, 2], np.int32)
    w = np.array([0.5, 0.25, 0.75])
    dense = np.zeros((4, 4))
    dense[ef, et] = w
    dense += dense.T
    import scipy.sparse as sp
    csr = sp.csr_matrix(dense)
    g = ConosGraph(names=names, vertices=np.arange(4, dtype=np.int32), edge_from=ef, edge_to=et, weight=w,
                   type=np.ones(3, np.int32), adj_p=csr.indptr.astype(np.int64), adj_i=csr.indices.astype(np.int32),
                   adj_x=csr.data)
    cell_ids = ["c0", "c1", "c2", "c3"]
    m = g.scanpy_matrices(cell_ids)
    perm = [1, 3, 0, 2]
    want = dense[np.ix_(perm, perm)]
    c = m["graph_connectivities"]
    got = sp.csc_matrix((c["data"], c["indices"], c["indptr"]), shape=tuple(c["shape"])).toarray()
    assert np.array_equal(got, want) and list(c["shape"]) == [4, 4]
    assert all(np.all(np.diff(c["indices"][c["indptr"][j]:c["indptr"][j + 1]]) > 0) for j in range(4))   # dgCMatrix order
    d = m["graph_distances"]
    assert np.array_equal(d["data"], 1.0 - c["data"]) and np.array_equal(d["indices"], c["indices"])
    assert np.array_equal(g.scanpy_matrices()["graph_connectivities"]["data"], sp.csc_matrix(dense).data)
    with pytest.raises(KeyError, match="not a vertex"):
        g.scanpy_matrices(["c0", "nope"])
    with pytest.raises(ValueError, match="extension"):
        g.save_for_scanpy("/tmp/x.hdf5")
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="h5py"):
            g.save_for_scanpy("/tmp/x.h5")
