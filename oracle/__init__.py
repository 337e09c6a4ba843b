"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference algorithm (C + Python) and the golden-vector generator."""
