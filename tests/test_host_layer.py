"""Host-side mirror of the reference's classes (safe_mpc_b200/{env_model,controller,utils,cost_definition}.py): the
parts that do not need a GPU (construction, name maps, error behaviour, pure-numpy predicates) and -- marked gpu -- the
same classes driving the engine against the flat ``Engine`` calls."""
import numpy as np
import pytest

from safe_mpc_b200 import abi
from safe_mpc_b200.parser import Parameters, default_args
from safe_mpc_b200.env_model import AdamModel
from safe_mpc_b200.utils import get_controller
from safe_mpc_b200 import controller as C
from safe_mpc_b200.cost_definition import ReachTargetEXT, ReachTargetNLS, ZeroCost
from tests.common import start_states, rollout_guess


def _model(batch=4, **over):
    args = default_args(**over)
    params = Parameters(args, 'z1', rti=True)
    params.N = args['horizon']
    return AdamModel(params, batch=batch), params


def test_get_controller_keys_and_errors():
    model, _ = _model(controller='st', horizon=12)
    names = {'naive': C.NaiveController, 'zerovel': C.TerminalZeroVelocity, 'st': C.STController, 'htwa': C.HTWAController,
             'receding': C.RecedingController, 'real_receding': C.RealReceding, 'constraint_everywhere': C.ControllerSafeSetEverywhere}
    for k, cls in names.items():
        assert type(get_controller(k, model)) is cls
    for bad in ('stwa', 'parallel', 'nope'):                 # utils.py:64-75 has no such keys
        with pytest.raises(ValueError, match='not available'):
            get_controller(bad, model)
    # the classes the reference defines without a CLI key exist and are wired to their engine state machines
    assert issubclass(C.ParallelController, C.RecedingController) and C.ParallelController.engine_name == 'parallel'
    assert C.STWAController.engine_name == 'stwa'
    from safe_mpc_b200.problem import CONTROLLERS
    from safe_mpc_b200 import abi
    assert CONTROLLERS['parallel'] == (abi.CTRL['parallel'], abi.NN_PARALLEL, False)


def test_model_attributes_and_predicates():
    model, params = _model(batch=3, joint_bounds_margin=5.0)
    assert (model.nq, model.nx, model.nu) == (5, 10, 5)
    # env_model.py:115-121: the model bounds are widened by q_margin percent of the range
    d = model.data
    assert np.allclose(model.x_max - model.x_min, d.bounds_diff * (1 + 2 * 0.05))
    x = np.vstack([0.5 * (model.x_min + model.x_max), model.x_max + 2 * params.tol_x, model.x_max + 0.5 * params.tol_x])
    assert model.checkStateBounds(x).tolist() == [True, False, True]
    tau = np.vstack([model.tau_max, model.tau_max + 1e-3])
    assert model.checkTorqueBounds(tau).tolist() == [True, False]
    model.reset_seed()
    assert np.abs(model.torque_noise).max() == 0.0         # control_noise = 0
    model.params.control_noise = 2.0
    model.reset_seed()
    expect = np.random.default_rng(1).normal(np.zeros(5), model.tau_max * 0.02, size=5)
    assert np.array_equal(model.torque_noise[1], expect)    # default_rng(test index), env_model.py:196,330-331
    model.update_randomized_dynamics(noise_percent=10.0, seed=0)
    rel = np.abs(model.plant_inertial[:, :, 0] / d.inertial[None, :, 0] - 1)
    assert 0 < rel.max() <= 0.1 + 1e-12                     # lumped masses move by at most the requested percentage


def test_controller_before_build_and_costs():
    model, _ = _model(controller='htwa', horizon=10)
    ctrl = get_controller('htwa', model)
    assert ctrl.N == 10 and ctrl.engine_name == 'htwa' and ctrl.cost_kind == 'ext'
    with pytest.raises(ValueError, match='not built'):
        ctrl.solve(np.zeros((4, 10)))
    ReachTargetNLS(model, 1e2, 5e-3).set_solver_cost(ctrl)
    assert ctrl.cost_kind == 'nls'
    ReachTargetEXT(model).set_solver_cost(ctrl)
    assert ctrl.cost_kind == 'ext'
    bk = C.SafeBackupController(model)
    assert bk.cost_kind == 'zero'
    ZeroCost(model).set_solver_cost(bk)
    assert bk.cost_kind == 'zero' and bk.engine_name == 'backup'
    assert ctrl.time_fields[-1] == 'time_tot'


@pytest.mark.gpu
def test_controller_classes_drive_the_engine_like_flat_calls():
    from safe_mpc_b200.engine import Engine
    from safe_mpc_b200.problem import build_problem
    B, N = 8, 12
    model, params = _model(batch=B, controller='receding', horizon=N)
    ctrl = get_controller('receding', model)
    ReachTargetEXT(model, params.Q_weight, params.R_weight).set_solver_cost(ctrl)
    ctrl.build_controller()
    prob, keep = build_problem(params, 'receding', cost='ext', N=N, model=model.data)
    eng = Engine(prob, B, 0)
    x0 = start_states(B, seed=5)
    xg, ug = rollout_guess(x0, N, params.dt, seed=6)
    ctrl.setGuess(xg, ug); ctrl.reset_controller()
    eng.set_guess(xg, ug); eng.reset_controller()
    x_c, x_e = x0.copy(), x0.copy()
    for _ in range(4):
        u_c, ab_c = ctrl.step(x_c)
        u_e, ab_e = eng.controller_step(x_e)
        assert np.array_equal(u_c, u_e) and np.array_equal(ab_c, ab_e.astype(bool))
        assert np.array_equal(ctrl.r, eng.get_state(abi.STATE_R)) and np.array_equal(ctrl.fails, eng.get_state(abi.STATE_FAILS))
        x_c, _ = model.integrate(x_c, u_c)
        x_e, _ = eng.plant_step(x_e, u_e)
        assert np.array_equal(x_c, x_e)
    assert ctrl.x_temp.shape == (B, N + 1, 10) and ctrl.getTime().shape == (7,)
    assert model.jointToEE(x_c).shape == (B, 3) and model.checkCollision(x_c).shape == (B,)
    assert ctrl.checkSafeConstraints(x_c).shape == (B,)


@pytest.mark.gpu
def test_mpc_script_runs_and_writes_the_reference_result_file(tmp_path, monkeypatch):
    import pickle
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location('mpc_script', os.path.join(os.path.dirname(__file__), '..', 'scripts', 'mpc.py'))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    monkeypatch.setattr(mod.Parameters, '__init__', _short_run(mod.Parameters.__init__, str(tmp_path) + '/', 40))
    # without a guess file the script fails like the reference does on its missing pickle (mpc.py:80) ...
    with pytest.raises(FileNotFoundError, match='guess_acados'):
        mod.main(['-c', 'htwa', '--horizon', '15', '--back_hor', '15', '--batch', '6'])
    # ... unless it is told to generate the warm starts itself
    n_coll = mod.main(['-c', 'htwa', '--horizon', '15', '--back_hor', '15', '--batch', '6', '--generate-guess'])
    files = [f for f in os.listdir(tmp_path) if f.endswith('_mpc.pkl')]
    assert len(files) == 1 and files[0].startswith('z1_htwa_use_netTrue_15hor_10sm_noise_0.0_control_noise0.0')
    data = pickle.load(open(tmp_path / files[0], 'rb'))
    assert set(data) == {'x', 'u', 'r', 'conv_idx', 'collisions_idx', 'unconv_idx', 'viable_idx', 'x_viable'}     # mpc.py:307-315
    assert data['x'].shape == (6, 41, 10) and data['u'].shape == (6, 40, 5)
    groups = data['conv_idx'] + data['collisions_idx'] + data['unconv_idx'] + data['viable_idx']
    assert sorted(groups) == list(range(6)) and n_coll == len(data['collisions_idx'])


def _short_run(orig, data_dir, n_steps):
    def init(self, *a, **k):
        orig(self, *a, **k)
        self.DATA_DIR = data_dir
        self.n_steps = n_steps
    return init


def test_generate_urdf_noise_script_and_model_reload(tmp_path, monkeypatch):
    import os
    """scripts/generate_urdf_noise.py writes one perturbed URDF per test and level with the reference's seeding schedule
    (default_rng(0) for the first level, default_rng(test_num) afterwards: generate_urdf_noise.py:32-36, utils.py:19-24), and a file read
    back gives the inertial parameters the in-memory draw gives."""
    import importlib.util
    import shutil
    from safe_mpc_b200.robot_model import nominal_link_inertials, randomized_link_inertials
    from safe_mpc_b200.urdf import URDF
    spec = importlib.util.spec_from_file_location('gen_noise', os.path.join(os.path.dirname(__file__), '..', 'scripts', 'generate_urdf_noise.py'))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    orig = mod.Parameters.__init__

    def init(self, *a, **k):
        orig(self, *a, **k)
        dst = tmp_path / 'z1.urdf'
        if not dst.exists():
            shutil.copy(self.robot_urdf, dst)
        self.robot_urdf, self.test_num = str(dst), 4
    monkeypatch.setattr(mod.Parameters, '__init__', init)
    files = mod.main(['--noises', '5.0', '10.0'])
    assert len(files) == 8 and all(os.path.isfile(f) for f in files)
    assert os.path.basename(files[0]) == 'z1_randomizednoise5.0_0.urdf' and os.path.basename(files[-1]) == 'z1_randomizednoise10.0_3.urdf'
    nominal = nominal_link_inertials(URDF.from_xml_file(str(tmp_path / 'z1.urdf')))
    for level, seed, fs in ((5.0, 0, files[:4]), (10.0, 4, files[4:])):
        want = randomized_link_inertials(nominal, level, level, level, 4, seed=seed)
        for i, f in enumerate(fs):
            got = nominal_link_inertials(URDF.from_xml_file(f))
            assert np.array_equal(got['mass'], want['mass'][i]) and np.array_equal(got['com'], want['com'][i])
            assert np.array_equal(got['inertia6'], want['inertia6'][i])
