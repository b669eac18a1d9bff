class HeteroData:  # placeholder: only needed so data_loader.py imports; not used by the oracle
    pass
