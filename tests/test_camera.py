"""Host camera helper vs the reference Camera (inputs only need to agree closely: they are fed verbatim to every implementation)."""
import numpy as np
import pytest

from oracle import oraclebind, refbind
from urho3d_b200 import camera

pytestmark = pytest.mark.skipif(not refbind.available(), reason="oracle/_ref/libu3dref.so not built")


@pytest.mark.parametrize("ortho", [False, True])
def test_camera_matches_reference(ortho):
    ref = refbind.Ref(0)
    rs = np.random.RandomState(0)
    for _ in range(20):
        pos = rs.uniform(-300, 300, 3)
        yaw, pitch = rs.uniform(0, 360), rs.uniform(-60, 60)
        fov, aspect, near, far = rs.uniform(20, 100), rs.uniform(0.5, 2.5), rs.uniform(0.05, 2.0), rs.uniform(50, 2000)
        size = rs.uniform(10, 500)
        v = camera.make_view(pos, yaw, pitch, fov, aspect, near, far, ortho=ortho, ortho_size=size)
        ref.camera_set(pos, camera.quat_from_euler(pitch, yaw, 0.0), fov, aspect, near, far, ortho=ortho, ortho_size=size)
        view, proj, planes, absn, nf, rev = ref.camera_get()
        assert np.allclose(view, v.view, rtol=1e-5, atol=2e-4)
        assert np.allclose(proj, v.proj, rtol=1e-6, atol=1e-7)
        assert np.allclose(planes, v.planes, rtol=1e-4, atol=2e-2)
        assert np.array_equal(absn, np.abs(planes[:, :3]))
    ref.close()


def test_view_proj_matches_reference_and_oracle_bitwise():
    ref = refbind.Ref(0)
    ob = oraclebind.OracleOB()
    ref.ob_set_size(64, 64)
    ob.set_size(64, 64)
    rs = np.random.RandomState(1)
    for _ in range(20):
        v = camera.make_view(rs.uniform(-300, 300, 3), rs.uniform(0, 360), rs.uniform(-60, 60), 45.0, 1.3, 0.1, 1000.0)
        ref.set_view_raw(v)
        ob.set_view(v)
        vp = camera.view_proj(v.proj, v.view)
        assert np.array_equal(ref.viewproj(), vp)
        assert np.array_equal(ob.viewproj(), vp)
    ref.close()
    ob.close()
