from .metrics import Metrics
from .nano_beir_evaluator import BaseEvaluator, Dataset, NanoBEIREvaluator, SparseSearch
from .ranker import Ranker, rank

__all__ = ['Metrics', 'Ranker', 'rank', 'SparseSearch', 'BaseEvaluator', 'NanoBEIREvaluator', 'Dataset']
