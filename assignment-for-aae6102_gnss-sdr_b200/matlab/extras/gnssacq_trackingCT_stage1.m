function [TckResultCT, CN0_Eph, countinx] = gnssacq_trackingCT_stage1(file, signal, track, Acquired)
% The FIRST (1 ms) stage of acqtckpos/trackingCT.m only (its lines 22-212): the conventional DLL/PLL loop of
% every acquired satellite runs on the GPU (gnssacq_mex 'track' -> gnssacq_track: one thread-block cluster per
% channel, no host round trip per millisecond); C/N0 and the bit-edge index are computed here as the reference does.
% NOT a replacement of trackingCT: the reference goes on (lines 214-533: bit-edge re-run, 40 s of 10 ms
% integrations) and returns the TckResultCT that naviDecode_updated consumes; this function returns the
% 1 ms records of stage 1.  It therefore has its own name and lives in matlab/extras/, which must NOT be put on the
% path in front of acqtckpos/ (only matlab/ itself, for acquisition.m, is): SDR_main.m:38 keeps calling the
% reference's trackingCT.  Call it explicitly where the stage-1 loop is wanted.
% Field note: TckResultCT(sv).codedelay(i) here is the plain cumulative sum of the DLL corrections
% (AcqCodeDelay + sum over THIS satellite's first i delays).  trackingCT.m:161 writes sum(delayValue(1:Index)) with a
% LINEAR index into the n_sv x n_ms matrix, which equals that only for a single satellite; the deviation is deliberate.
n_ms = track.msToProcessCT_1ms;
N = signal.Sample;
bps = file.dataPrecision * file.dataType;
cfg = struct('fs_hz', signal.Fs, 'if_hz', signal.IF, 'code_hz', signal.codeFreqBasis, 'samples_per_ms', N, ...
             'data_type', file.dataType, 'data_precision', file.dataPrecision, 'noncoh_blocks', 1, 'prn', 1);
fseek(file.fid, file.skip * N * bps, 'bof');                         % one read instead of an fread per ms
if file.dataPrecision == 2
    seg = fread(file.fid, (n_ms + 3) * N * file.dataType, 'int16=>int16');
else
    seg = fread(file.fid, (n_ms + 3) * N * file.dataType, 'int8=>int8');
end
gnssacq_mex('track_load', seg, cfg);
n_sv = length(Acquired.sv);
ch = zeros(7, n_sv);
for k = 1:n_sv                                                      % trackingCT.m:42-60
    ch(:, k) = [Acquired.sv(k); 0; N - Acquired.codedelay(k) + 1; Acquired.fineFreq(k); 0; signal.codeFreqBasis; 0];
end
loops = [track.DLLBW track.DLLDamp track.DLLGain track.PLLBW track.PLLDamp track.PLLGain track.CorrelatorSpacing];
rec = gnssacq_mex('track', ch, cfg, loops, n_ms);                   % 14 x n_ms x n_sv
countinx = zeros(1, n_sv);
K = 20;
for k = 1:n_sv
    sv = Acquired.sv(k);
    r = rec(:, :, k);
    delay = r(14, :) - N * track.pdi;
    TckResultCT(sv).P_i = r(1, :);  TckResultCT(sv).P_q = r(2, :);
    TckResultCT(sv).E_i = r(3, :);  TckResultCT(sv).E_q = r(4, :);
    TckResultCT(sv).L_i = r(5, :);  TckResultCT(sv).L_q = r(6, :);
    TckResultCT(sv).PLLdiscri = r(7, :);  TckResultCT(sv).DLLdiscri = r(8, :);
    TckResultCT(sv).codedelay = Acquired.codedelay(k) + cumsum(delay);
    TckResultCT(sv).remChip = r(9, :);  TckResultCT(sv).codeFreq = r(10, :);
    TckResultCT(sv).carrierFreq = r(11, :);  TckResultCT(sv).remPhase = r(12, :);
    TckResultCT(sv).numSample = r(14, :);  TckResultCT(sv).delayValue = delay;
    TckResultCT(sv).absoluteSample = (r(13, :) + file.skip * N) * bps;
    TckResultCT(sv).codedelay2 = mod(TckResultCT(sv).absoluteSample / bps, signal.Fs * signal.ms);
    Zk = r(1, :).^2 + r(2, :).^2;                                   % trackingCT.m:121-133
    for b = 1:floor(n_ms / K)
        z = Zk((b - 1) * K + 1 : b * K);
        NA2 = sqrt(mean(z)^2 - var(z));
        varIQ = 0.5 * (mean(z) - NA2);
        CN0_Eph(b, k) = abs(10 * log10(1 / (1 * signal.ms * track.pdi) * NA2 / (2 * varIQ)));
    end
    P = TckResultCT(sv).P_i;                                        % trackingCT.m:176-212
    for i = max(7, 600):length(P) - 18
        if all(sign(P(i-6:i-1)) ~= sign(P(i))) && all(sign(P(i+1:i+17)) == sign(P(i)))
            countinx(k) = mod(i, 20) - 1;
            break
        end
    end
end
end
