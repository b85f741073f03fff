"""CPU oracle for the myDetection post-processing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mydetection_b200/`` imports this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may use it, and only as the checker or
as the timed CPU baseline -- never as the product path.

What it is: a restatement, on the CPU, of the reference's algorithm for
decode -> threshold -> top-k -> per-class NMS, the IoU primitives and the ATSS
assignment.  Elementwise float arithmetic is written with
``torch`` CPU operators because the reference's own arithmetic *is* torch CPU
arithmetic (same ``exp`` / ``sigmoid`` kernels => the restatement can be pinned
bit-for-bit against the reference); loop-heavy integer/geometry work (greedy NMS,
rotated polygon clipping) is plain C in ``oracle/*.c`` built by ``oracle/build.py``.

Pinning status (see DESIGN.md section "Oracle"):
  * decode (YOLOv3 / FCOS / FCOS2 / RAPiD / RetinaNet / YOLOv5), bboxes_iou,
    post_process, per-class AABB NMS, ATSS assignment: PINNED -- checked
    bit-exactly against outputs of the unmodified reference imported from
    /root/reference (fixtures in tests/golden/, generator
    tests/golden/make_golden.py) and against torchvision.ops.nms.
  * rotated IoU / nms_rotbb: PARITY UNPINNED.  The reference rasterises polygons
    with pycocotools (utils/bbox_ops.py:94-96), which is neither vendored nor
    installed and cannot be fetched.  The oracle is exact fp64 convex-polygon
    clipping wrapped in nms_rotbb's own control flow (utils/bbox_ops.py:276-306);
    the control flow is pinned against the reference with a stub IoU, the IoU
    values themselves are not.
"""
from . import decode, postprocess, iou, atss  # noqa: F401
