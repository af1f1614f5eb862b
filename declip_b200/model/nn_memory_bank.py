"""NNMemoryBankModule — mirror of prototype/model/utils/nnclr_modules/{memory_bank,nn_memory_bank}.py.

Same semantics (FIFO bank of `size` keys, initialised with normalised Gaussian noise, lookups see the bank BEFORE
the update of the same call, wrap-around writes only the tail and resets the pointer, memory_bank.py:71-87, returns
the UN-normalised bank rows of the top-k most cosine-similar keys), but the bank lives in HBM (the reference keeps
128 MiB on the CPU and uploads it on every call, nn_memory_bank.py:54) as fp32 rows [size, dim] plus a bf16
normalised shadow that feeds the similarity GEMM.  Only topk = 1 (every reference config) is built."""
import torch
from torch import nn

from .. import functions as F_


class NNMemoryBankModule(nn.Module):
    def __init__(self, size=2 ** 16, topk=1):
        super().__init__()
        if size < 0:
            raise ValueError('Illegal memory bank size %d, must be non-negative.' % size)
        if topk != 1:
            raise NotImplementedError("declip_b200: nn_topk != 1 is not used by any reference config")
        self.size = size
        self.topk = topk
        self.bank = None        # fp32 [size, dim]  (row i == reference bank[:, i])
        self.bank16 = None      # bf16 F.normalize(bank, dim=1)
        self.bank_ptr = 0

    @torch.no_grad()
    def _init_memory_bank(self, dim, device):
        bank = torch.randn(dim, self.size)                                       # memory_bank.py:66-67
        bank = torch.nn.functional.normalize(bank, dim=0)
        self.load_bank(bank, device)

    @torch.no_grad()
    def load_bank(self, bank_dim_by_size, device, ptr=0):
        """Install a bank given in the reference layout [dim, size] (parity tests, checkpoint migration)."""
        self.bank = bank_dim_by_size.t().contiguous().float().to(device)
        self.bank16 = F_.normalize_rows_bf16(self.bank)
        self.bank_ptr = int(ptr)

    @torch.no_grad()
    def _dequeue_and_enqueue(self, batch):
        bs = batch.shape[0]
        ptr = self.bank_ptr
        if ptr + bs >= self.size:                                                # memory_bank.py:81-84
            n = self.size - ptr
            self.bank[ptr:] = batch[:n]
            self.bank16[ptr:] = F_.normalize_rows_bf16(batch[:n]) if n > 0 else self.bank16[ptr:]
            self.bank_ptr = 0
        else:
            self.bank[ptr:ptr + bs] = batch
            self.bank16[ptr:ptr + bs] = F_.normalize_rows_bf16(batch)
            self.bank_ptr = ptr + bs

    @torch.no_grad()
    def forward(self, output, update=False):
        if self.size == 0:
            return [output]
        output = output.detach().float()
        if self.bank is None or self.bank.device != output.device:
            if self.bank is None:
                self._init_memory_bank(output.shape[1], output.device)
            else:
                self.bank, self.bank16 = self.bank.to(output.device), self.bank16.to(output.device)
        nearest, idx = F_.nn_lookup(output, self.bank, self.bank16)              # uses the pre-update bank
        self.last_index = idx
        if update:
            self._dequeue_and_enqueue(output)
        return [nearest]
