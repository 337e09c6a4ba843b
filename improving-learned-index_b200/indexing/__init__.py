"""Impact quantization and the doc-major collection text format (host side of K1)."""
