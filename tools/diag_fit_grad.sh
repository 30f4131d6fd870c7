for c in 0 1; do echo "== LEMO_GEMM_CLUSTER=$c"; LEMO_GEMM_CLUSTER=$c timeout 300 python - <<'PY'
import sys; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
import torch, numpy as np
import test_gpu_fit as t
from gpu_common import rel, rel_q, oracle_ctx
from oracle import ref_loops as rl
from lemo_b200 import _lib
for conv in ('pair','wt'):
    _lib.call('lemo_debug_set_conv_tc', t._CONV_MODES[conv])
    T,S=119,2
    c32,c64=oracle_ctx(torch.float32),oracle_ctx(torch.float64)
    fit=t._fitter(S,T,use_cuda_graph=True)
    probs=[t._problem(s,T,c32) for s in range(S)]
    for s,(init,mrec,con) in enumerate(probs): fit.set_sequence(s,init,mrec,con)
    fit.run(n_iters=1); st=fit.state()
    for s,(init,mrec,con) in enumerate(probs):
        tr64=[]; rl.fit_temp(init,mrec,con,c64,n_iters=1,faithful=False,trace=tr64)
        sl=slice(s*T,(s+1)*T)
        for k in ('g_transl','g_rot6d','g_other'):
            a=st[k][sl]; b=tr64[0][k]
            print(conv,s,k,'max %.2e q999 %.2e q99 %.2e med %.2e'%(rel(a,b),rel_q(a,b,.999),rel_q(a,b,.99),rel_q(a,b,.5)))
PY
done
