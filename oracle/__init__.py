"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's caption-generation path
(thenoobychocobo/gpt2-image-captioning: src/models.py, src/database/faiss_store.py
and the third-party arithmetic it calls: transformers GPT2LMHeadModel 4.57.3,
torch nn.Linear / nn.TransformerEncoder 2.9.1, faiss-cpu 1.13.1 IndexFlatIP).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import anything from here, and only as the checker
(never as the thing measured as the product, never as a fallback).  The product
package `gpt2_image_captioning_b200` must not import this package.

Parity status: the reference has NO tests, golden vectors or fixtures for this
path, so the pin is "outputs of the reference itself run
here": `oracle/ref_harness.py` imports the unmodified reference from
/root/reference (build container only) and `tests/golden/make_golden.py` commits
its token ids / logits as fixtures; `tests/test_oracle.py` checks this restatement
against those fixtures token for token.
"""
