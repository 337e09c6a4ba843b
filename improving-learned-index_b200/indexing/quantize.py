"""Linear 8-bit impact quantization of a doc-major collection file — drop-in for the
reference's src/deep_impact/indexing/quantize.py (quantize :13-14, find_max_value :17-24,
quantize_file :27-47, CLI :50-58).

Text parsing and formatting stay in Python; the arithmetic (global max, and
``int(score * (255 / max))`` in float64) runs in the CUDA kernels of csrc/build.cuh (K1).
"""
from __future__ import annotations

import argparse
import logging
from pathlib import Path
from typing import List, Optional, Tuple, Union

import numpy as np

from .. import engine
from ..utils.defaults import IMPACT_SCORE_QUANTIZATION_BITS

logger = logging.getLogger('quantize')
_CHUNK_LINES = 200_000
_INT32_MAX = 2 ** 31 - 1


def quantize(value: float, scale: float) -> int:
    """Scalar form, kept for API compatibility (quantize.py:13-14). The batch path is GPU-side."""
    return int(value * scale)


def _parse_chunk(lines: List[str]) -> Tuple[List[List[str]], np.ndarray]:
    """lines -> per-line term lists + one flat float64 array of all scores (file order)."""
    terms_per_line, scores = [], []
    for line in lines:
        terms = []
        for pair in line.strip().split(', '):
            term, score = pair.strip().split(': ')      # blank line -> ValueError, as in the reference
            terms.append(term)
            scores.append(float(score))
        terms_per_line.append(terms)
    return terms_per_line, np.asarray(scores, dtype=np.float64)


def _chunks(path):
    with open(path, 'r', encoding='utf-8') as f:
        chunk = []
        for line in f:
            chunk.append(line)
            if len(chunk) == _CHUNK_LINES:
                yield chunk
                chunk = []
        if chunk:
            yield chunk


def find_max_value(input_file_path: Union[str, Path]):
    max_val = 0
    for chunk in _chunks(input_file_path):
        _, scores = _parse_chunk(chunk)
        if scores.size:
            max_val = max(max_val, engine.find_max(scores))
    return max_val


def quantize_file(input_file_path: Union[str, Path], output_file_path: Union[str, Path],
                  max_val: Optional[float] = None):
    if max_val is None:
        max_val = find_max_value(input_file_path)
        logger.info(f'Found max value: {max_val}')
    else:
        logger.info(f'Using given max value: {max_val}')
    scale_check = ((1 << IMPACT_SCORE_QUANTIZATION_BITS) - 1) / max_val   # ZeroDivisionError like the reference

    with open(output_file_path, 'w', encoding='utf-8') as out:
        for chunk in _chunks(input_file_path):
            terms_per_line, scores = _parse_chunk(chunk)
            if not np.isfinite(scores * scale_check).all():
                raise OverflowError('cannot convert a non-finite impact to an integer')
            values = engine.quantize(scores, max_val)
            if values.size and (np.abs(values.astype(np.int64)) >= _INT32_MAX).any():
                raise OverflowError('quantized impact does not fit 32 bits')
            values = values.tolist()
            pos = 0
            for terms in terms_per_line:
                kept = [f'{t}: {v}' for t, v in zip(terms, values[pos:pos + len(terms)]) if v > 0]
                pos += len(terms)
                out.write(', '.join(kept) + '\n')


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description='Quantize a DeepImpact collection.')
    parser.add_argument('-i', '--input_file_path', type=Path, required=True)
    parser.add_argument('-o', '--output_file_path', type=Path, required=True)
    parser.add_argument('-m', '--max_val', type=float, default=None)
    quantize_file(**vars(parser.parse_args()))
