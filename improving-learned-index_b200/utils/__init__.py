"""On-disk format constants and the query / qrels / run-file wire formats."""
