#!/usr/bin/env python
"""Wall-clock of the text pipeline (collection text -> quantized text -> index files -> loaded index) on a
synthetic 100K-document collection; run on the GPU box. Checks the output against the CPU oracle."""
import sys, time, tempfile, hashlib
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from improving_learned_index_b200 import synthetic as syn, quantize_file, InvertedIndexCreator, InvertedIndex
from oracle import oracle

n_docs = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
c = syn.make_collection(n_docs, vocab_size=30522, draws_per_doc=120, seed=5)
tmp = Path(tempfile.mkdtemp())
raw = tmp / "collection.index"
raw.write_text(''.join(l + '\n' for l in c.lines()), encoding='utf-8')
t0 = time.time(); quantize_file(raw, tmp / "q"); t1 = time.time()
InvertedIndexCreator(tmp / "q", tmp / "index").run(); t2 = time.time()
index = InvertedIndex(tmp / "index"); t3 = time.time()
q = oracle.quantize(c.impacts); keep = q > 0
used = sorted(set(c.term_ids[keep].tolist()))
remap = np.full(c.vocab_size, -1, dtype=np.int64); remap[used] = np.arange(len(used))
doc_of = np.repeat(np.arange(c.n_docs), np.diff(c.doc_offsets.astype(np.int64)))
offs = np.zeros(c.n_docs + 1, dtype=np.uint64); offs[1:] = np.cumsum(np.bincount(doc_of[keep], minlength=c.n_docs))
toff, docs, imps = oracle.invert(remap[c.term_ids[keep]], q[keep], offs, len(used))
dat, idx = oracle.serialize(toff, docs, imps)
ok = (dat.tobytes() == (tmp / "index" / "inverted_index.dat").read_bytes()
      and idx.tobytes() == (tmp / "index" / "inverted_index.idx").read_bytes())
print(f"{n_docs} docs, {int(keep.sum())} postings, text {raw.stat().st_size/1e6:.0f} MB: quantize_file {t1-t0:.2f}s, "
      f"InvertedIndexCreator {t2-t1:.2f}s, InvertedIndex load {t3-t2:.2f}s, files byte-identical to oracle: {ok}")
