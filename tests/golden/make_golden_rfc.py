"""Regenerates tests/golden/rfc_zk_vector1.json from the reference's own known-answer test
`test_zk_rfc_testvector1` (rust/runtime/zk/tests/zk.rs:228-558: a 3-term circuit over GF(2^128),
rate 4, nreq 6, block_enc 128, transcript seed "test", with the byte arrays the C++ prover produced:
commitment root, sumcheck proof, Ligero proof).  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_rfc.py

Only the byte arrays and the parameters of that test are recorded; the witness and the coin stream are
rebuilt by tests/fixtures.py:load_rfc_vector from the description in the test (w = [1, embed(5), embed(6),
(embed(5)+embed(6))*x]; a RandomEngine whose every bytes(n) call returns 02 00 ... 00)."""
import json
import os
import re

SRC = "/root/reference/rust/runtime/zk/tests/zk.rs"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = open(SRC).read()
    body = src[src.index("fn test_zk_rfc_testvector1"):src.index("fn test_zk_triple_zero_zero_zero")]

    def arr(name):
        m = re.search(r"let %s: &\[u8\] = &\[(.*?)\];" % name, body, re.S)
        return bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", m.group(1))).hex()

    cfg = re.search(r"rateinv: (\d+),\s*nreq: (\d+),\s*block_enc: (\d+)", body)
    out = dict(source="rust/runtime/zk/tests/zk.rs:228-558 (test_zk_rfc_testvector1)",
               field_id=4, rate=int(cfg.group(1)), nreq=int(cfg.group(2)), block_enc=int(cfg.group(3)),
               transcript_seed="test", subfield_scalars=[5, 6],
               circuit=arr("circuit_bytes"), sumcheck_proof=arr("expected_sc_proof"),
               commitment=arr("expected_com"), ligero_proof=arr("expected_com_proof"))
    json.dump(out, open(os.path.join(HERE, "rfc_zk_vector1.json"), "w"), indent=1)
    print("wrote rfc_zk_vector1.json: circuit %d B, proof %d B" % (
        len(out["circuit"]) // 2, (len(out["commitment"]) + len(out["sumcheck_proof"]) + len(out["ligero_proof"])) // 2))


if __name__ == "__main__":
    main()
