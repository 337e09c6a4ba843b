"""On-disk format contract of the reference index (src/utils/defaults.py:22-37 there).

Only the constants the inverted-index path needs; nothing here touches torch or the GPU.
"""
import struct

INVERTED_INDEX_VOCAB = 'vocab.txt'            # one term per line, line number = term id
INVERTED_INDEX_INDEX = 'inverted_index.idx'   # per term: start byte, end byte into the .dat (2 x u64)
INVERTED_INDEX_DATA = 'inverted_index.dat'    # per posting: u32 docid + u8 impact, unpadded

IMPACT_SCORE_QUANTIZATION_BITS = 8
IMPACT_SCORE_FORMAT, IMPACT_SCORE_BYTES = 'B', 1
DOC_ID_FORMAT, DOC_ID_BYTES = 'I', 4
LOC_FORMAT, LOC_BYTES = 'Q', 8

DOC_SCORE_BLOCK_FORMAT = DOC_ID_FORMAT + IMPACT_SCORE_FORMAT
DOC_SCORE_BLOCK_BYTES = DOC_ID_BYTES + IMPACT_SCORE_BYTES
LOC_BLOCK_FORMAT = LOC_FORMAT * 2
LOC_BLOCK_BYTES = LOC_BYTES * 2

COLLECTION_TYPES = ['msmarco', 'beir']
MAX_IMPACT = (1 << IMPACT_SCORE_QUANTIZATION_BITS) - 1

assert struct.calcsize('<' + DOC_SCORE_BLOCK_FORMAT) == DOC_SCORE_BLOCK_BYTES
assert struct.calcsize('<' + LOC_BLOCK_FORMAT) == LOC_BLOCK_BYTES
