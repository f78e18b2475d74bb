"""ORACLE / TEST INFRASTRUCTURE -- literal CPU restatement of the VAR identification of the reference's driver,
README.md:116-130:

    for i = PN+1:num_train
        for j = 1:PN,  AA(i-PN, n*(j-1)+1 : n*j) = ad_acc(i-j, :);  end
        BB(i-PN, :) = ad_acc(i, :);
    end
    PARA = (AA'*AA)\\AA'*BB;          % MATLAB precedence: ((AA'*AA)\\AA') * BB
    A1 = (PARA(1:n, :))';  A2 = (PARA(1+n : 2n, :))';

PARITY UNPINNED by the reference (MATLAB, no golden vectors, cannot run here): pinned by recovery of a known model and
agreement with the benchmark generator's own identification (tests/test_oracle_varid.py).
Only tests/, bench.py's baseline leg and __graft_entry__.smoke() may import this module."""
import numpy as np


def identify(ad_acc: np.ndarray, PN: int = 2):
    """ad_acc: (num_train, n).  Returns A of shape (PN, n, n) with A[j-1] = A_j."""
    K, n = ad_acc.shape
    AA = np.zeros((K - PN, PN * n))
    BB = np.zeros((K - PN, n))
    for i in range(PN, K):                       # README.md:123  i = PN+1:num_train (1-based)
        for j in range(1, PN + 1):               # :124
            AA[i - PN, n * (j - 1):n * j] = ad_acc[i - j]
        BB[i - PN] = ad_acc[i]                   # :127
    PARA = np.linalg.solve(AA.T @ AA, AA.T) @ BB     # :130, MATLAB's left-to-right evaluation
    return np.stack([PARA[n * j:n * (j + 1)].T for j in range(PN)])
