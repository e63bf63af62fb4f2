function [TckResultCT, CN0_Eph, countinx] = trackingCT(file, signal, track, Acquired)
% Drop-in for the first (1 ms) stage of acqtckpos/trackingCT.m (lines 22-212): the conventional DLL/PLL loop of
% every acquired satellite runs on the GPU (gnssacq_mex 'track' -> gnssacq_track: one thread-block cluster per
% channel, no host round trip per millisecond); C/N0 and the bit-edge index are computed here exactly as the
% reference does.  The later stages of the reference (lines 214-533) are not replaced.
%   addpath('<repo>/assignment-for-aae6102_gnss-sdr_b200/matlab', '-begin');
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
gnssacq_mex(seg, cfg, 'track_load');
n_sv = length(Acquired.sv);
ch = zeros(7, n_sv);
for k = 1:n_sv                                                      % trackingCT.m:42-60
    ch(:, k) = [Acquired.sv(k); 0; N - Acquired.codedelay(k) + 1; Acquired.fineFreq(k); 0; signal.codeFreqBasis; 0];
end
loops = [track.DLLBW track.DLLDamp track.DLLGain track.PLLBW track.PLLDamp track.PLLGain track.CorrelatorSpacing];
rec = gnssacq_mex(ch, cfg, loops, n_ms, 'track');                   % 14 x n_ms x n_sv
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
