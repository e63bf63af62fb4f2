function Acquired = acquisition(file, signal, acq)
% GPU drop-in for acqtckpos/acquisition.m (same signature, same result fields).
% The coarse search runs in libgnssacq.so through gnssacq_mex; this wrapper only reads the raw
% bytes (a MATLAB file id cannot cross into C), maps the three structs onto the library
% config, and assembles the result struct and the progress lines callers expect.
% Put this directory ahead of acqtckpos on the path:  addpath(<this dir>, '-begin')

cfg = struct('fs_hz', signal.Fs, 'if_hz', signal.IF, 'code_hz', signal.codeFreqBasis, ...
    'samples_per_ms', signal.Sample, 'data_type', file.dataType, 'data_precision', file.dataPrecision, ...
    'freq_min_hz', acq.freqMin, 'freq_step_hz', acq.freqStep, 'freq_num', acq.freqNum, ...
    'noncoh_blocks', acq.datalen, 'coh_ms', 1, 'snr_threshold_db', 12, 'prn', 1:32, 'n_gpus', 1);
if isfield(acq, 'nGpus'), cfg.n_gpus = acq.nGpus; end    % PRN-major shards over several GPUs (gnssacq_search_multi)

kinds = {'int8=>int8', 'int16=>int16'};
bytesPerMs = signal.Sample * file.dataPrecision * file.dataType;
fseek(file.fid, file.skip * bytesPerMs, 'bof');
raw = fread(file.fid, signal.Sample * file.dataType * acq.datalen, kinds{file.dataPrecision});

fprintf('Acquiring... \n ');
for svindex = 1:32, svindex, end       %#ok<NOPRT> acquisition.m:48 echoes the loop index; kept for identical console output
rows = gnssacq_mex('search', raw, cfg);            % n_prn x [prn acquired code_phase bin doppler peak noise snr]
hit = rows(rows(:, 2) == 1, :);

Acquired = struct('sv', [], 'SNR', [], 'Doppler', [], 'codedelay', [], 'fineFreq', []);
if isempty(hit)
    fprintf('No satellites acquired. Check parameter settings ... \n\n ');
    return
end
Acquired.sv        = hit(:, 1).';
Acquired.SNR       = hit(:, 8).';
Acquired.Doppler   = hit(:, 5).';
Acquired.codedelay = hit(:, 3).';
for k = 1:size(hit, 1)
    fprintf(' SV[%2d] SNR = %2.2f, Code phase = %5d, Raw Doppler = %5d \n ', ...
        hit(k, 1), hit(k, 8), hit(k, 3), hit(k, 5));
end
% Fine-frequency refinement, also on the GPU: the gateway takes the (L+1) ms block and the acquired
% (PRN, code phase) pairs and returns the absolute carrier frequency of each.
fprintf('Now refining Doppler freq... \n ');
fseek(file.fid, file.skip * bytesPerMs, 'bof');
longraw = fread(file.fid, signal.Sample * file.dataType * (acq.L + 1), kinds{file.dataPrecision});
Acquired.fineFreq = gnssacq_mex('fine', longraw, cfg, acq.L, Acquired.sv, Acquired.codedelay);
for k = 1:numel(Acquired.sv)
    fprintf(' SV[%2d] SNR = %2.2f, Code phase = %5d, Raw Doppler = %5d, Fine Doppler = %5f \n ', ...
        Acquired.sv(k), Acquired.SNR(k), Acquired.codedelay(k), Acquired.Doppler(k), Acquired.fineFreq(k) - signal.IF);
end
end
