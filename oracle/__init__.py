"""CPU oracle for the CSMPN hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package.  Nothing under clifford-group-equivariant-simplicial-message-passing-networks_b200/ imports it.

Contents
  layers_ref.py  plain-PyTorch (CPU, fp32 or fp64, autograd-differentiable) restatement of the reference's
                 algebra and layers, written from the formulas with explicit index/sign tables.
  lift_ref.py    pure-Python restatement of the simplicial lifting (csmpn/data/modules/{utils,simplicial_data}.py)
                 over a small SimplexTree stand-in.
  refshim.py     sys.modules stand-ins (torch_geometric, torch_scatter, gudhi, ...) that let the UNMODIFIED
                 reference be imported from /root/reference in the build container; used by
                 tests/golden/make_golden.py to generate the committed fixtures and by the optional
                 "reference present" tests.  /root/reference does not exist on the GPU box.

Pinning status: the reference ships no tests and no golden vectors.  layers_ref.py is
pinned against outputs of the reference's own code run in the build container (tests/golden/*.pt, generated
by tests/golden/make_golden.py; re-checked live by tests/test_oracle_golden.py when /root/reference is
mounted).  lift_ref.py is pinned the same way for everything except gudhi's traversal orders (gudhi is not
installable here): for those parts parity is UNPINNED against real gudhi and pinned only against the
documented SimplexTree semantics; motion's ManualTransform is literal and fully pinned.
"""
