"""The residency rules of pybader_b200.session on the CPU, with a recording stand-in for the
device handle (the real thing is exercised on the GPU by tests/test_gpu_boundary.py):
which host arrays are uploaded, copied on the device or taken as resident, and into which slot.
"""
import numpy as np
import pytest

from pybader_b200 import session
from pybader_b200.engine import LABELS_ATOMS, LABELS_BADER, RHO_CHARGE, RHO_REFERENCE, RHO_SPIN


class FakeEngine:
    def __init__(self, shape, device=0):
        self.shape, self.log = tuple(shape), []
        self.rho, self.lab = {}, {}

    def upload_density(self, which, rho):
        self.log.append(('upload_density', which))
        self.rho[which] = np.array(rho, dtype=np.float64)

    def copy_density(self, dst, src):
        self.log.append(('copy_density', dst, src))
        self.rho[dst] = self.rho[src].copy()

    def upload_labels(self, which, labels):
        self.log.append(('upload_labels', which))
        self.lab[which] = np.array(labels, dtype=np.int64)

    def clear_labels(self, which):
        self.log.append(('clear_labels', which))
        self.lab[which] = np.zeros(self.shape, dtype=np.int64)

    def download_labels(self, which, dtype=np.int32, out=None):
        self.log.append(('download_labels', which))
        if out is None:
            return self.lab[which].astype(dtype)
        out[...] = self.lab[which]
        return out

    def close(self):
        pass


@pytest.fixture()
def sess(monkeypatch):
    session.close_all()
    monkeypatch.setattr(session, 'Engine', FakeEngine)
    s = session.get((4, 5, 6))
    yield s
    session._sessions.clear()


def test_reference_is_always_slot_zero(sess):
    rng = np.random.default_rng(0)
    a, b = rng.random((4, 5, 6)), rng.random((4, 5, 6))
    assert sess.reference(a) == RHO_REFERENCE and sess.engine.log == [('upload_density', RHO_REFERENCE)]
    assert sess.reference(a) == RHO_REFERENCE and len(sess.engine.log) == 1          # resident
    assert sess.reference(a.copy()) == RHO_REFERENCE and len(sess.engine.log) == 1   # equal content
    # b becomes resident in the CHARGE slot (e.g. through charge_sum) ...
    assert sess.density_slot(b, prefer=sess.free_density_slot()) == RHO_CHARGE
    # ... and is then used as the reference: copied on the device into slot 0, never assumed
    sess.engine.log.clear()
    assert sess.reference(b) == RHO_REFERENCE
    assert sess.engine.log == [('copy_density', RHO_REFERENCE, RHO_CHARGE)]
    np.testing.assert_array_equal(sess.engine.rho[RHO_REFERENCE], b)
    # back to a: slot 0 no longer holds it -> uploaded again
    sess.engine.log.clear()
    assert sess.reference(a) == RHO_REFERENCE and sess.engine.log == [('upload_density', RHO_REFERENCE)]
    np.testing.assert_array_equal(sess.engine.rho[RHO_REFERENCE], a)
    assert sess.uploads['density'] == 3


def test_in_place_edit_of_a_density_is_seen(sess):
    a = np.random.default_rng(1).random((4, 5, 6))
    sess.reference(a)
    a[3, 4, 5] += 1e-12                      # one voxel, far from any stride a sample would take
    sess.engine.log.clear()
    sess.reference(a)
    assert sess.engine.log == [('upload_density', RHO_REFERENCE)]
    np.testing.assert_array_equal(sess.engine.rho[RHO_REFERENCE], a)


def test_density_slots_fill_reference_charge_spin(sess):
    rng = np.random.default_rng(2)
    a, b, c = (rng.random((4, 5, 6)) for _ in range(3))
    assert sess.density_slot(a, prefer=sess.free_density_slot()) == RHO_REFERENCE
    assert sess.density_slot(b, prefer=sess.free_density_slot()) == RHO_CHARGE
    assert sess.density_slot(c, prefer=sess.free_density_slot()) == RHO_SPIN
    assert sess.density_slot(b, prefer=RHO_SPIN) == RHO_CHARGE          # found where it lives
    assert [e for e in sess.engine.log if e[0] == 'upload_density'] == \
        [('upload_density', RHO_REFERENCE), ('upload_density', RHO_CHARGE), ('upload_density', RHO_SPIN)]


def test_labels_fresh_zeros_are_cleared_not_uploaded(sess):
    z = np.zeros((4, 5, 6), dtype=np.int32)
    assert sess.label_slot(z, force=LABELS_BADER) == LABELS_BADER
    assert sess.engine.log == [('clear_labels', LABELS_BADER)] and sess.uploads['labels'] == 0
    lab = np.arange(120, dtype=np.int16).reshape(4, 5, 6) % 7
    assert sess.label_slot(lab, force=LABELS_BADER) == LABELS_BADER
    assert sess.engine.log[-1] == ('upload_labels', LABELS_BADER) and sess.uploads['labels'] == 1
    assert sess.label_slot(lab, force=LABELS_BADER) == LABELS_BADER and sess.uploads['labels'] == 1


def test_labels_found_in_either_slot_unless_forced(sess):
    lab = (np.arange(120).reshape(4, 5, 6) % 5).astype(np.int8)
    sess.engine.lab[LABELS_ATOMS] = lab.astype(np.int64)
    host = sess.labels_to_host(LABELS_ATOMS, np.int8)          # what assign_to_atoms hands out
    assert sess.label_slot(host) == LABELS_ATOMS                # refine / surface_distance find it there
    n = sess.uploads['labels']
    assert sess.label_slot(host, force=LABELS_BADER) == LABELS_BADER   # bader_calc needs it in slot 0
    assert sess.uploads['labels'] == n + 1
    # an edit of the array the engine handed out is a different array
    host[0, 0, 1] = 3 if host[0, 0, 1] != 3 else 4
    assert sess.label_slot(host) == LABELS_BADER and sess.uploads['labels'] == n + 2
    np.testing.assert_array_equal(sess.engine.lab[LABELS_BADER], host)


def test_one_grid_resident_at_a_time(monkeypatch):
    session.close_all()
    monkeypatch.setattr(session, 'Engine', FakeEngine)
    a = session.get((2, 2, 2))
    assert session.get((2, 2, 2)) is a
    b = session.get((3, 2, 2))
    assert b is not a and list(session._sessions) == [(3, 2, 2)]
    session._sessions.clear()
