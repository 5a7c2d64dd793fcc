from torch import Tensor
Adj = Tensor
OptTensor = Tensor
