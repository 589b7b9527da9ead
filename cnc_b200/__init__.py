"""cnc-b200: the data-parallel hot path of YihangChen-ee/CNC for NVIDIA B200 (sm_100a).

Host side = the reference's own Python operator surface; everything below it is hand-written CUDA behind the C ABI of
`include/cnc_b200.h` (`cnc_b200/lib/libcnc_b200.so`, built by `python -m cnc_b200.build`).  There is no CPU path.

    cnc_b200.gridencoder       GridEncoder, STE_binary                        (examples/radiance_fields/ngp.py:22-315)
    cnc_b200.field             NGPRadianceField_mygrid_2D3D                   (ngp.py:365-645)
    cnc_b200.nerfacc           OccGridEstimator, traverse_grids, rendering .. (vendored nerfacc 0.5.3)
    cnc_b200.render            render_image_with_occgrid[_test]               (examples/utils.py:83-489)
    cnc_b200.context_models    CNC_context_models                             (examples/utils_bpp_acc.py:193-999)
    cnc_b200.container         one-file bitstream container
    cnc_b200.trainer / .dp     the training step and its data-parallel exchange
    cnc_b200._gridencoder / .pack_and_align / .torchac     drop-ins for the reference's three native modules
"""
__version__ = "0.2.0"
