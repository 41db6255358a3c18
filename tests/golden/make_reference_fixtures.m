% make_reference_fixtures.m -- pins the numpy oracle (and through it the CUDA path) against the reference.
%
% Run with Matlab (or Octave with classdef support) from the repository root of THIS project:
%
%   >> addpath('/path/to/ska-pst-dsp-model/matlab');      % the reference checkout
%   >> run('tests/golden/make_reference_fixtures.m');
%
% It loads tests/golden/reference_inputs.mat (written by export_reference_inputs.py: the exact seeded inputs of
% tests/golden/*.npz), calls the reference's own functions in the call forms of test_data_pipeline.m:114,130-132
% and channelize.m:91,96, once on the single-precision inputs (the harness default, default_config.m:23) and once
% on the same inputs cast to double, and writes tests/golden/reference_outputs.mat (v7, readable by scipy).
% tests/test_golden.py::test_reference_outputs_pin_oracle then asserts oracle == reference (<= 1e-10 double,
% <= 1e-5 single) and test_cuda_matches_reference asserts the GPU result against the same arrays.

here = fileparts(mfilename('fullpath'));
if isempty(here); here = fullfile(pwd, 'tests', 'golden'); end
S = load(fullfile(here, 'reference_inputs.mat'));
names = fieldnames(S);
R = struct();
win = PFBWindow();
for i = 1:numel(names)
  c = S.(names{i});
  os_factor = struct('nu', double(c.nu), 'de', double(c.de));
  for prec = {'single', 'double'}
    x = cast(c.x, prec{1});
    switch strtrim(c.func)
      case 'polyphase_analysis'
        y = polyphase_analysis(x, c.filt, double(c.nchan), os_factor);
      case 'polyphase_analysis_padded'
        y = polyphase_analysis_padded(x, c.filt, double(c.nchan), os_factor);
      case 'polyphase_analysis_lowcbf'
        clear polyphase_analysis_lowcbf;                       % persistent first-call flag (:27-34)
        y = polyphase_analysis_lowcbf(x, c.filt, double(c.nchan), os_factor);
      case 'polyphase_synthesis'
        factory = win.lookup('tukey');
        taper = factory(double(c.nfft), double(c.overlap));
        deripple = struct('apply_deripple', 1, 'filter_coeff', c.filt);
        % test_data_pipeline.m:130-132: input_fully_spans_Nyquist_zone = 1, sample_offset = 1
        y = polyphase_synthesis(x, 1, double(c.nfft), os_factor, deripple, 1, double(c.overlap), taper, @identity_taper);
      otherwise
        error('unknown function %s', c.func);
    end
    R.([names{i} '_' prec{1}]) = y;
  end
end
save(fullfile(here, 'reference_outputs.mat'), '-struct', 'R', '-v7');
fprintf('wrote %s\n', fullfile(here, 'reference_outputs.mat'));
