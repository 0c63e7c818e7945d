"""TEST INFRASTRUCTURE ONLY - a stand-in package with the reference's module layout and call sites
(`ultralytics.utils.{tal,loss,ops}`, `ultralytics.nn.modules.head`), whose bodies delegate to the
oracle. The GPU box has no /root/reference, so `kuzushiji_vision_b200.install()` is exercised against
this package there: same attribute names, same `from .tal import TaskAlignedAssigner` binding in
`loss`, same `Detect.forward -> _inference` and `v8DetectionLoss.__call__` call flow
(yolov12/ultralytics/utils/loss.py:206-260, nn/modules/head.py:64-74)."""
