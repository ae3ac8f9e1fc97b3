"""Device string columns (ops/strings.py, csrc/strings.cu): host reference paths
here; the same checks run against the kernels under the gpu marker."""
import numpy as np
import pytest
import torch

from bigslice_b200.frame import hashing
from bigslice_b200.ops import strings as S

WORDS = ["the quick  brown fox", "", "jumps\tover the\nlazy dog", "  x  ", "naïve café",
         "a", "the the the", "Zeta\x0bend"]
URLS = ["https://Site12.Example.com/page/1?q=2", "http://user:pw@Host.ORG:8080/x", "example.net/path",
        "ftp://a.b.c", "", "https://UPPER.case/#frag"]


def _check_all(dev):
    col = S.from_objects(WORDS, device=dev)
    assert len(col) == len(WORDS) and col.to_objects().tolist() == WORDS
    # murmur3 over bytes, bit-exact with the host reference
    for seed in (0, 0x9acb0442):
        want = hashing.hash_objects(np.array(WORDS, dtype=object), seed)
        assert torch.equal(S.hash32(col, seed).cpu(), want)
    # tokeniser == str.split() on the same delimiters
    toks, parent = S.split(col)
    want_toks, want_parent = [], []
    for i, w in enumerate(WORDS):
        for t in w.split():
            want_toks.append(t)
            want_parent.append(i)
    assert toks.to_objects().tolist() == want_toks
    assert parent.cpu().tolist() == want_parent
    # gather
    idx = torch.tensor([6, 0, 0, 3])
    assert S.take(col, idx).to_objects().tolist() == [WORDS[i] for i in idx.tolist()]
    # dictionary ids: equal strings <-> equal ids, representatives verified
    ids = S.key64(toks)
    d_ids, d_strs = S.dictionary(toks, ids)
    assert len(d_ids) == len(set(want_toks))
    assert sorted(d_strs.to_objects().tolist()) == sorted(set(want_toks))
    assert S.verify_dictionary(toks, ids, d_ids, d_strs)
    # a forged collision is caught
    bad = ids.clone()
    bad[0] = bad[1]
    if want_toks[0] != want_toks[1]:
        bd_ids, bd_strs = S.dictionary(toks, bad)
        assert not S.verify_dictionary(toks, bad, bd_ids, bd_strs)
    # URL hosts
    u = S.from_objects(URLS, device=dev)
    assert S.url_domain(u).to_objects().tolist() == ["site12.example.com", "host.org:8080", "example.net",
                                                     "a.b.c", "", "upper.case"]


def test_strings_host():
    _check_all(None)


@pytest.mark.gpu
def test_strings_device_kernels():
    from bigslice_b200.ops import gpu
    before = gpu.launches()
    _check_all("cuda")
    assert gpu.launches() - before >= 8
    # bulk: 1M tokens hashed on the device equal the host hashes
    rng = np.random.default_rng(0)
    words = np.array(["w%d" % i for i in rng.integers(0, 50000, 200000)], dtype=object)
    col = S.from_objects(words, device="cuda")
    assert torch.equal(S.hash32(col, 0).cpu(), hashing.hash_objects(words, 0))


def test_csr_columns_flow_through_combinators(local_session):
    """StringColumn (CSR) columns through Map/Flatmap/Filter/Reduce/shuffles on the
    CPU executor — the same column type the device pipelines carry."""
    import bigslice_b200 as bs
    from bigslice_b200 import text
    lines = np.array(["a b a", "c  b", "", "a"] * 50, dtype=object)
    D = text.IdDictionary()

    @bs.Func
    def job():
        s = bs.Map(bs.Const(3, lines), lambda c: (S.from_objects(c),), out=[bs.string])   # host CSR
        toks = text.Tokenize(s)
        toks = bs.Filter(toks, lambda c: torch.as_tensor([t != "c" for t in c.tolist()]))
        keyed = bs.Map(toks, lambda c: (c, torch.ones(len(c), dtype=torch.int64)), out=[bs.string, bs.int64])
        return bs.Reduce(keyed, bs.ops.add)          # string-keyed shuffle + combine
    rows = dict(local_session.run(job).rows())
    assert rows == {"a": 150, "b": 100}

    @bs.Func
    def ids():
        s = bs.Map(bs.Const(2, lines), lambda c: (S.from_objects(c),), out=[bs.string])
        return bs.Reduce(text.StringIds(text.Tokenize(s), D), bs.ops.add)
    got = local_session.run(ids).rows()
    names = D.lookup([k for k, _ in got])
    assert {names[k]: v for k, v in got} == {"a": 150, "b": 100, "c": 50}
