function make_reference_golden(reference_root, golden_dir)
% MAKE_REFERENCE_GOLDEN  run the UNMODIFIED reference classes (Lib/CFS_FANUC.m, Lib/PSGCFS_FANUC.m of
% JessicaLeu-code/MotionPlanning_5D_m) on the golden configurations of this repository and write their results in the flat
% CFSB format, so that the CPU oracle and the CUDA path can be pinned against real MATLAB + quadprog output.
%
%   make_reference_golden('/path/to/MotionPlanning_5D_m', '/path/to/repo/tests/golden')
%
% reads   <golden_dir>/fixtures/<case>.bin        (python tests/golden/export_fixtures.py)
% writes  <golden_dir>/reference/<case>_ref.bin   { u, x_, cost_all, iter_O, total_iter, matlab_release }
% then    python -m pytest tests/test_reference_golden.py [-m gpu]   compares oracle / GPU with these files.
%
% Needs MATLAB with the Optimization Toolbox (quadprog) and, for PSGCFS, nothing else: normrnd is SHADOWED by a shim that
% replays the fixture's noise columns in call order (the reference class stays untouched; Statistics Toolbox not required).
% DO NOT put <repo>/matlab's CFS_FANUC.m before Lib/ on the path for this script: it must run the reference's own classes.
here = fileparts(mfilename('fullpath'));
addpath(fullfile(reference_root, 'Lib'), fullfile(reference_root, 'Lib', 'functions'), fullfile(reference_root, 'Lib', 'M16iB'), ...
        fullfile(reference_root, 'Lib', '200i'), fullfile(reference_root, 'Lib', '2L'));
w = which('CFS_FANUC');
if ~isempty(strfind(w, here)), error('cfs:golden', 'CFS_FANUC resolves to the drop-in (%s): remove %s from the path', w, here); end
shim = tempname;  mkdir(shim);
fid = fopen(fullfile(shim, 'normrnd.m'), 'w');
fprintf(fid, 'function r = normrnd(varargin)\n%% replays the fixture''s draws (PSGCFS_FANUC.m:109), one column per call\nglobal CFS_NOISE CFS_NOISE_K\nCFS_NOISE_K = CFS_NOISE_K + 1;\nr = CFS_NOISE(:, CFS_NOISE_K);\nend\n');
fclose(fid);
names = {'M16iB', 'M200i', '2L'};
files = dir(fullfile(golden_dir, 'fixtures', '*.bin'));
if ~exist(fullfile(golden_dir, 'reference'), 'dir'), mkdir(fullfile(golden_dir, 'reference')); end
for f = 1:numel(files)
    fx = read_cfs_fixture(fullfile(files(f).folder, files(f).name));
    ROBOT = names{fx.robot_id + 1};
    robot = robotproperty2(ROBOT);
    obs = cell(1, size(fx.obs_l, 3));
    for j = 1:numel(obs)
        obs{j}.shape = 'cylinder';  obs{j}.A = eye(2);
        obs{j}.l = fx.obs_l(:, :, j);  obs{j}.D = fx.obs_D(j);  obs{j}.epsilon = fx.obs_epsilon(j);
    end
    H = fx.H;  nj = fx.njoint;
    sys_info = struct();                                             % main_FANUC.m:106-127
    sys_info.Aaug = fx.Aaug;  sys_info.Baug = fx.Baug;  sys_info.QQ = fx.QQ;  sys_info.ff = fx.ff(:);
    sys_info.Qaug = fx.QQ;  sys_info.paug = fx.ff(:);  sys_info.caug = fx.caug;  sys_info.robot = robot;
    sys_info.H = H;  sys_info.nstate = 2 * nj;  sys_info.njoint = nj;  sys_info.nu = nj;
    sys_info.xR = fx.x0(:);  sys_info.x_ = fx.x_(:);  sys_info.alpha = fx.alpha;
    if fx.has_lim, sys_info.lim = fx.lim(:); end
    sys_info.epsilon_O = fx.epsilon_O;  sys_info.MAX_O_ITER = fx.MAX_O_ITER;  sys_info.MAX_input = fx.MAX_input(:);
    if fx.solver_id == 1
        global CFS_NOISE CFS_NOISE_K %#ok<TLEV>
        CFS_NOISE = fx.noise;  CFS_NOISE_K = 0;
        addpath(shim);
        self = PSGCFS_FANUC(obs, sys_info, ROBOT);
        self = self.optimizer();
        rmpath(shim);
    else
        self = CFS_FANUC(obs, sys_info, ROBOT);
        self = self.optimizer();
    end
    out = struct('u', self.u(:), 'x_', self.x_(:), 'cost_all', self.eval.cost_all(:), 'iter_O', self.iter_O, ...
                 'total_iter', self.total_iter, 'matlab_release', double(version('-release')));
    [~, base] = fileparts(files(f).name);
    write_cfs_fixture(fullfile(golden_dir, 'reference', [base '_ref.bin']), out);
    fprintf('%s: iter_O = %d, final cost = %.10g\n', base, self.iter_O, self.eval.cost_all(end));
end
end
