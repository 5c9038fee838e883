import torch.utils.data


class ListDataset(torch.utils.data.Dataset):
    """Dataset over a list of elements, each turned into a sample by `load` (torchnet.dataset.ListDataset)."""

    def __init__(self, elem_list, load=lambda x: x, path=None):
        self.list, self.load = list(elem_list), load

    def __len__(self):
        return len(self.list)

    def __getitem__(self, idx):
        if idx < 0 or idx >= len(self):
            raise IndexError("index out of range")
        return self.load(self.list[idx])
