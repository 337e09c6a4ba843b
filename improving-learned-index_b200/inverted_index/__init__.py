from .create import InvertedIndexCreator
from .inverted_index import InvertedIndex

__all__ = ['InvertedIndex', 'InvertedIndexCreator']
