"""Linear 8-bit impact quantization of a doc-major collection file — drop-in for the
reference's src/deep_impact/indexing/quantize.py (quantize :13-14, find_max_value :17-24,
quantize_file :27-47, CLI :50-58).

Text parsing and formatting are host work (the library's C++ parser, collection_io.py, with the
reference's tokenisation rules); the arithmetic (global max, and ``int(score * (255 / max))`` in
float64) runs in the CUDA kernels of csrc/build.cuh (K1).
"""
from __future__ import annotations

import argparse
import logging
from pathlib import Path
from typing import Optional, Union

import numpy as np

from .. import collection_io, engine
from ..utils.defaults import IMPACT_SCORE_QUANTIZATION_BITS

logger = logging.getLogger('quantize')
_INT32_MAX = 2 ** 31 - 1


def quantize(value: float, scale: float) -> int:
    """Scalar form, kept for API compatibility (quantize.py:13-14). The batch path is GPU-side."""
    return int(value * scale)


def find_max_value(input_file_path: Union[str, Path]):
    """quantize.py:17-24: max over every score of the file, seeded with 0 (GPU reduction, K1)."""
    parsed = collection_io.parse_file(input_file_path, collection_io.SEQUENCE)
    try:
        return max(0, engine.find_max(parsed.scores)) if parsed.scores.size else 0
    finally:
        parsed.close()


def quantize_file(input_file_path: Union[str, Path], output_file_path: Union[str, Path],
                  max_val: Optional[float] = None):
    parsed = collection_io.parse_file(input_file_path, collection_io.SEQUENCE)   # ValueError on a blank line
    try:
        if max_val is None:
            max_val = max(0, engine.find_max(parsed.scores)) if parsed.scores.size else 0
            logger.info(f'Found max value: {max_val}')
        else:
            logger.info(f'Using given max value: {max_val}')
        scale = ((1 << IMPACT_SCORE_QUANTIZATION_BITS) - 1) / max_val        # ZeroDivisionError like the reference
        if parsed.scores.size and not np.isfinite(parsed.scores * scale).all():
            raise OverflowError('cannot convert a non-finite impact to an integer')
        values = engine.quantize(parsed.scores, max_val) if parsed.scores.size else np.zeros(0, dtype=np.int32)
        if values.size and (np.abs(values.astype(np.int64)) >= _INT32_MAX).any():
            raise OverflowError('quantized impact does not fit 32 bits')
        parsed.write_quantized(values, output_file_path)
    finally:
        parsed.close()


if __name__ == "__main__":
    parser = argparse.ArgumentParser(description='Quantize a DeepImpact collection.')
    parser.add_argument('-i', '--input_file_path', type=Path, required=True)
    parser.add_argument('-o', '--output_file_path', type=Path, required=True)
    parser.add_argument('-m', '--max_val', type=float, default=None)
    quantize_file(**vars(parser.parse_args()))
