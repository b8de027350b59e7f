"""CPU tests of the native snarkjs-JSON ingestion (csrc/keyjson.cu, host-only entry points of libzkr):
zkr_pkey_json_to_bin / zkr_witness_json_to_bin must produce exactly the bytes of the reference's
binarifyProvingKey / binarifyWitness (operator/src/utils/binarify.ts:10-207), checked against the oracle's
line-by-line restatement (oracle/binfmt.py) and the Python host mirror, on oracle-generated keys."""
import ctypes as C
import json
import random

import pytest

from oracle import binfmt as bf
from oracle import groth16 as g
from oracle.bn254 import Q, R
from simple_zk_rollups_b200 import _lib, binarify, synth

TOXIC = (1234567891011, 222222222222223, 3333333333333331, 44444444444447, 5555555555555557)


def _key(nc, npub, seed):
    r1, w = synth.generate(nc, npub, seed=seed)
    pk, vk, _ = g.setup(r1.to_dicts(), TOXIC)
    return pk, w


@pytest.fixture(scope="module")
def small():
    return _key(40, 3, 11)


def test_pkey_json_matches_reference_layout(small):
    pk, _ = small
    want = bf.binarify_proving_key(pk)
    j = bf.pk_to_json(pk)
    for text in (json.dumps(j), json.dumps(j, indent=2), json.dumps(j, separators=(",", ":"))):
        got = binarify.binarifyProvingKeyJson(text)
        assert got == want
    # the Python host mirror (dict input) and the native text path agree as well
    assert binarify.binarifyProvingKey(j) == want


def test_pkey_json_key_order_numbers_and_unknown_fields(small):
    pk, _ = small
    want = bf.binarify_proving_key(pk)
    j = bf.pk_to_json(pk)
    rnd = random.Random(5)
    keys = list(j)
    rnd.shuffle(keys)
    j2 = {k: j[k] for k in keys}
    j2["somethingElse"] = {"nested": [1, 2, {"x": "]}\\\""}], "s": "a,b]"}
    # polynomial objects with keys out of order: Object.keys still enumerates ascending
    j2["polsA"] = [dict(reversed(list(col.items()))) for col in j2["polsA"]]
    # nVars etc. as strings (stringifyBigInts output of a BigInt field) instead of JSON numbers
    j2["nVars"] = str(j2["nVars"])
    assert binarify.binarifyProvingKeyJson(json.dumps(j2)) == want


def test_pkey_json_infinity_and_unreduced_values(small):
    pk, _ = small
    j = bf.pk_to_json(pk)
    # snarkjs writes the affine zero as ["0","1","0"]; binarify drops z -> (0, R mod q)
    assert any(p == ["0", "1", "0"] for p in j["B1"]), "fixture should contain an infinity in B1"
    # values >= modulus are reduced by times(2^256).mod(p) (binarify.ts:82,89)
    j["A"][1][0] = str(int(j["A"][1][0]) + Q)
    j["polsB"][0] = {"0": str(R + 5)}
    pk2 = dict(pk)
    pk2["polsB"] = [dict(c) for c in pk["polsB"]]
    pk2["polsB"][0] = {0: 5}
    assert binarify.binarifyProvingKeyJson(json.dumps(j)) == bf.binarify_proving_key(pk2)


def test_pkey_json_roundtrip_through_parser(small):
    pk, _ = small
    blob = binarify.binarifyProvingKeyJson(json.dumps(bf.pk_to_json(pk)))
    back = bf.parse_proving_key(blob)
    for k in ("nVars", "nPublic", "domainSize", "polsA", "polsB", "A", "B1", "B2", "hExps"):
        assert back[k] == pk[k], k
    assert back["C"][pk["nPublic"] + 1:] == pk["C"][pk["nPublic"] + 1:]


def test_pkey_json_larger_key():
    pk, _ = _key(300, 4, 3)
    assert binarify.binarifyProvingKeyJson(json.dumps(bf.pk_to_json(pk))) == bf.binarify_proving_key(pk)


@pytest.mark.parametrize("mutate, needle", [
    (lambda j: j.pop("hExps"), "hExps missing"),
    (lambda j: j.pop("nPublic"), "missing"),
    (lambda j: j["A"].pop(), "A has"),
    (lambda j: j["A"].__setitem__(0, ["12x", "1", "1"]), "decimal"),
    (lambda j: j["A"].__setitem__(0, [str(1 << 256), "1", "1"]), "decimal"),
    (lambda j: j["C"].__setitem__(len(j["C"]) - 1, None), "null point"),
    (lambda j: j["polsA"].__setitem__(0, {"-1": "1"}), "row index"),
    (lambda j: j["B2"].__setitem__(0, [["1"], ["1", "0"], ["1", "0"]]), "two coefficients"),
])
def test_pkey_json_rejects_malformed(small, mutate, needle):
    pk, _ = small
    j = bf.pk_to_json(pk)
    mutate(j)
    with pytest.raises(_lib.ZkrError) as e:
        binarify.binarifyProvingKeyJson(json.dumps(j))
    assert e.value.code == -2 and needle in str(e.value), str(e.value)


def test_pkey_json_truncated_text(small):
    pk, _ = small
    text = json.dumps(bf.pk_to_json(pk))
    for cut in (0, 1, len(text) // 3, len(text) - 1):
        with pytest.raises(_lib.ZkrError):
            binarify.binarifyProvingKeyJson(text[:cut])


def test_witness_json_matches_reference_layout(small):
    _, w = small
    want = bf.binarify_witness(w)
    assert binarify.binarifyWitnessJson(json.dumps([str(x) for x in w])) == want
    assert binarify.binarifyWitnessJson(json.dumps([str(x) for x in w], indent=1)) == want
    assert binarify.binarifyWitness(w) == want
    assert binarify.binarifyWitnessJson("[]") == b""
    assert binarify.binarifyWitnessJson("[1, 2, \"3\"]") == bf.binarify_witness([1, 2, 3])
    # edge values: 0, r-1, 2^256-1 are written as they are (binarify.ts:18-26)
    edge = [1, 0, R - 1, (1 << 256) - 1]
    assert binarify.binarifyWitnessJson(json.dumps([str(x) for x in edge])) == bf.binarify_witness(edge)
    with pytest.raises(_lib.ZkrError):
        binarify.binarifyWitnessJson("[\"1\", \"abc\"]")
    with pytest.raises(_lib.ZkrError):
        binarify.binarifyWitnessJson("[\"%d\"]" % (1 << 256))


def test_null_arguments():
    L = _lib.lib()
    out, n = C.c_void_p(), C.c_size_t()
    assert L.zkr_pkey_json_to_bin(None, 0, C.byref(out), C.byref(n)) == -1
    assert L.zkr_witness_json_to_bin(b"[]", 2, None, C.byref(n)) == -1
    L.zkr_buf_free(None)


def test_pkey_json_fabricated_rollup_shape_matches_host_mirror():
    """A few thousand signals of random field values in the tx.circom key shape (73 public signals, polsC present,
    1-3 non-zeros per column): the native converter equals the line-by-line Python mirror of binarify.ts."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from bench_keyjson import fabricate
    j = fabricate(3000, 73, seed=5)
    text = json.dumps(j)
    got = binarify.binarifyProvingKeyJson(text)
    assert got == binarify.binarifyProvingKey(json.loads(text))
    n, l, m = j["nVars"], j["nPublic"], j["domainSize"]
    assert int.from_bytes(got[0:4], "little") == n and int.from_bytes(got[4:8], "little") == l
    assert int.from_bytes(got[36:40], "little") + 64 * m == len(got)          # pPointsHExps + hExps = end (binarify.ts:199-204)
