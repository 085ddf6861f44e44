"""Host logic of the multi-GPU runner (dualsphysics_b200/run.py) that needs no GPU: the PART schedule (the same rule as the C++
driver, JDsOutputTime::GetNextTime) and the run times taken from the case with the command-line overrides."""
import pytest

import helpers as H
from dualsphysics_b200 import run


def test_next_part_time_is_the_first_multiple_after_t():
    assert run.next_part_time(0.0, 0.005) == 0.005
    assert run.next_part_time(0.005, 0.005) == 0.01            # a PART saved exactly on the mark: the next mark
    assert run.next_part_time(0.00586, 0.005) == 0.01
    assert run.next_part_time(0.0149, 0.005) == 0.015
    k, t = 0, 0.0
    while t < 0.02:                                            # the schedule of tests/test_host_driver_frozen.py: 5 PARTs with dt=1/1024
        t += 1.0 / 1024.0
        if t >= (run.next_part_time(0.0, 0.005) if k == 0 else tn):
            k += 1
            tn = run.next_part_time(t, 0.005)
    assert k == 4


def test_run_times_from_case_and_overrides(tmp_path):
    H.write_case(tmp_path, dp=0.05, params={"TimeMax": 0.3, "TimeOut": 0.02})
    case = str(tmp_path / "Case")
    assert run.run_times(case, ["-symplectic"]) == (0.3, 0.02, 0)
    assert run.run_times(case + ".xml", ["-tmax:0.1", "-TOUT:0.05", "-nsteps:7"]) == (0.1, 0.05, 7)
    with pytest.raises(ValueError):
        run.run_times(case, ["-tout:0"])


def test_usage_without_arguments(capsys):
    assert run.main([]) == 1 and "torch.distributed.run" in capsys.readouterr().out
