"""CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/darray_oracle.py header).

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU legs.
"""
