"""The CPU oracle and the scene generators still produce the outputs pinned in tests/golden/oracle_pins.json
(tests/golden/make_oracle_pins.py).  These pins guard the oracle against drift; they are not reference-derived
(the reference has no golden vectors for this path: "parity unpinned")."""
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))


def test_oracle_outputs_match_the_committed_pins():
    spec = importlib.util.spec_from_file_location("make_oracle_pins", os.path.join(HERE, "golden", "make_oracle_pins.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(HERE, "golden", "oracle_pins.json")))
    got = mod.compute()
    assert sorted(got) == sorted(want)
    for k in want:
        assert got[k] == want[k], k
