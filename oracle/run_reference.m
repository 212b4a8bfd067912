% run_reference.m - runs the UNMODIFIED function files of JiaoXianjun/multi-rtl-sdr-calibration on the inputs written by
% oracle/export_fixtures.py and saves every output, so that tests/test_reference_run.py can pin oracle/gsmcal_oracle.py
% (and through it the CUDA path) to the reference itself.  GNU Octave and MATLAB both run it; only core functions are needed
% (filter, fft, interp1, toeplitz, kron): no Signal / Communications / Instrument Control toolbox.
%
%   cd tests/golden/reference_run
%   octave --no-gui --eval "run_reference('/path/to/multi-rtl-sdr-calibration')"
%
% The current directory must hold the case_*.mat / planted_*.mat files and gsm_chn_filter_{8x,4x}.mat (chn_filter_*.m load
% the latter from the current directory).  Outputs: out_<case>.mat next to the inputs.
%
% Call order and arguments of the capture cases are those of gsm_sync_demod.m:107-124.  Two inputs are passed in rather
% than generated, because the reference generates them with toolboxes that are closed source / differ in Octave:
%   coef (gsm_sync_demod.m:34 fir1(46, 200e3/fs))  and  tpl (gsm_sync_demod.m:38 gsm_SCH_training_sequence_gen(8)).
function run_reference(ref_dir)
  if nargin < 1, ref_dir = '.'; end
  addpath(ref_dir);
  is_octave = exist('OCTAVE_VERSION', 'builtin') ~= 0;
  files = dir('case_*.mat');
  for k = 1:numel(files)
    in = load(files(k).name);
    name = files(k).name(6:end-4);
    s = double(in.raw);                                            % fread(..., 'uint8') returns double (gsm_sync_demod.m:96)
    r0 = raw2iq(s);                                                % gsm_sync_demod.m:107
    r = filter(in.coef, 1, r0);                                    % :110
    dec = r(1:64:end);
    [position, snr] = FCCH_coarse_position(dec, 8);                % :117
    first = dec(1:min(numel(dec), ceil(23*1250/8)));
    [mv_flag, mv_idx, mv_avg, mv_snr] = move_fft_snr_runtime_avg(first, 160, 16, 10);
    [FCCH_pos, r1, sppm1, cppm1] = FCCH_fine_correction(r, position, in.osr, in.carrier_freq);      % :118
    [pos_info, r2, sppm2] = SCH_corr_rate_correction(r1, FCCH_pos, in.tpl, in.osr);                 % :119
    [r3, cppm2] = carrier_correct_post_SCH(r2, pos_info, in.osr, in.carrier_freq);                  % :120
    total_sampling_ppm = total_ppm_calculation([sppm1 sppm2]);     % :123
    total_carrier_ppm = total_ppm_calculation([cppm1 cppm2]);      % :124
    c8 = chn_filter_8x_4x(r0(1:20000));
    c4 = chn_filter_4x(r0(1:20000));
    % whole streams are large: keep a strided sample of each plus its length (the test compares at these indices)
    stride = 997;
    out = struct('position', position, 'snr', snr, 'mv_flag', double(mv_flag), 'mv_idx', mv_idx, 'mv_avg', mv_avg, 'mv_snr', mv_snr, ...
                 'FCCH_pos', FCCH_pos, 'sppm1', sppm1, 'cppm1', cppm1, 'pos_info', pos_info, 'sppm2', sppm2, 'cppm2', cppm2, ...
                 'total_sampling_ppm', total_sampling_ppm, 'total_carrier_ppm', total_carrier_ppm, ...
                 'r0_head', r0(1:4096), 'r_len', numel(r), 'r_s', r(1:stride:end), ...
                 'r1_len', numel(r1), 'r1_s', r1(1:stride:end), 'r2_len', numel(r2), 'r2_s', r2(1:stride:end), ...
                 'r3_len', numel(r3), 'r3_s', r3(1:stride:end), 'chn8', c8, 'chn4', c4, 'stride', stride);
    save_struct(['out_' name '.mat'], out, is_octave);
    printf_line(['done ' name]);
  end
  files = dir('planted_*.mat');
  for k = 1:numel(files)
    in = load(files(k).name);
    name = files(k).name(9:end-4);
    stride = 997;
    if isfield(in, 'base_position')
      [FCCH_pos, r1, sppm1, cppm1] = FCCH_fine_correction(in.s, in.base_position, in.osr, in.carrier_freq);
      out = struct('FCCH_pos', FCCH_pos, 'sppm1', sppm1, 'cppm1', cppm1, 'r1_len', numel(r1), 'r1_s', r1(1:stride:end), 'stride', stride);
    else
      [pos_info, r2, sppm2] = SCH_corr_rate_correction(in.s, in.FCCH_pos, in.tpl, in.osr);
      out = struct('pos_info', pos_info, 'sppm2', sppm2, 'r2_len', numel(r2), 'r2_s', r2(1:stride:end), 'stride', stride);
    end
    save_struct(['out_' name '.mat'], out, is_octave);
    printf_line(['done ' name]);
  end
end

function save_struct(fname, out, is_octave)
  if is_octave
    save('-mat7-binary', fname, '-struct', 'out');
  else
    save(fname, '-struct', 'out', '-v7');
  end
end

function printf_line(t)
  disp(t);
end
