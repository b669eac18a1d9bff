class DataLoader:  # placeholder: test.py's loader is not on the hot path
    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        self.dataset = dataset

    def __iter__(self):
        return iter(self.dataset)
