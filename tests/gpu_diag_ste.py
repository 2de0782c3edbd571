"""GPU diagnostic: per-parameter errors of the tensor-core backward against the bf16-network autograd."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(__file__)); sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import test_gpu_backward as T


def main():
    orig_assert = None
    # re-run the test body but collect instead of asserting
    import types
    src = open(os.path.join(os.path.dirname(__file__), "test_gpu_backward.py")).read()
    body = src[src.index("def test_tensor_core_backward_matches_bf16_network_autograd"):]
    body = body.replace('        assert err <= 3e-2, f"{k}: relative error {err:.3e} (|ref| {float(ref[k].norm()):.3e})"',
                        '        print(f"{k:42s} rel {err:.3e} |ref| {float(ref[k].norm()):.3e} |got| {float(got[k].norm()):.3e}")')
    body += '''
    W = torch.cat([P["transient_rgb.0.weight"], P["transient_sigma.weight"], P["transient_flow_fw.0.weight"], P["transient_flow_bw.0.weight"]], 0)
    db = torch.cat([got["transient_rgb.0.bias"], got["transient_sigma.bias"], got["transient_flow_fw.0.bias"], got["transient_flow_bw.0.bias"]], 0)
    ident = W.to(torch.bfloat16).float().T @ db
    print("identity W^T db vs got final.bias:", float((ident - got["transient_xyz_encoding_final.bias"]).norm() / ident.norm()),
          " vs ref:", float((ident - ref["transient_xyz_encoding_final.bias"]).norm() / ident.norm()))
'''
    ns = dict(T.__dict__)
    exec(body.replace("@pytest.mark.parametrize", "#"), ns)
    ns["test_tensor_core_backward_matches_bf16_network_autograd"]("main", 11, 37)


main()
