"""Writes tests/golden/state_dict_keys.json: parameter/buffer names and shapes of the reference
networks (constructed from /root/reference), the contract for checkpoint compatibility."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _reference  # noqa: E402


def main():
    pu, ssg, msg = _reference.partsize()
    p2u, dg, am, mm = _reference.highway()
    nets = {
        "partsize.pointnet2_sem_seg.get_model(13)": ssg.get_model(13),
        "partsize.pointnet2_sem_seg_msg.get_model(5)": msg.get_model(5),
        "highway.DGCNN.DGCNN(5, 20)": dg.DGCNN(5, 20),
        "highway.model.PointNet2(5)": mm.PointNet2(5),
        "highway.model.EnhancedPointNet2(5)": mm.EnhancedPointNet2(5),
        "highway.model.BridgeStructureLoss()": mm.BridgeStructureLoss(),
    }
    out = {k: [[n, list(t.shape)] for n, t in v.state_dict().items()] for k, v in nets.items()}
    json.dump(out, open(os.path.join(HERE, "state_dict_keys.json"), "w"))
    print({k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
