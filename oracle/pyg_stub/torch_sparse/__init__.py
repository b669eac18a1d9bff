class SparseTensor:  # placeholder type (periodGATconv.py:7 only uses it in isinstance checks)
    pass
