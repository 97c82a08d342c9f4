"""CPU oracle for the E2ENet hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (torch fp32 / numpy), the algorithms of the
reference's per-patch hot path:

  * oracle.network  -- depth-shift + (1,3,3) conv + InstanceNorm + LeakyReLU and the
                       UNet++-style DSFF fusion grid (reference:
                       e2enet/network_architecture/unetpp_d.py:38-111, 447-488)
  * oracle.masking  -- kernel-granular Masking init / apply / prune / regrow
                       (reference: e2enet/training/network_training/sparselearning/
                       core_channel.py:141-169, 290-317, 427-434, 556-611, 647-666, 721-739)
  * oracle.window   -- sliding-window tiling, Gaussian importance map and the
                       weighted accumulate (reference:
                       e2enet/network_architecture/neural_network.py:244-426, 500-565)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it, and only as the checker or the timed CPU baseline.  The product
package (e2enet_medical_b200) never imports it and has no CPU fallback.

Parity pinning: every function here is checked in tests/test_oracle_golden.py against
golden vectors produced by running the UNMODIFIED reference modules in the build
container (tests/golden/make_golden.py is the generator, the .npz/.json files under
tests/golden/ are its committed outputs), plus the reference's own known-answer
vectors for _compute_steps_for_sliding_window
(tests/test_steps_for_sliding_window_prediction.py:96-163).
"""
