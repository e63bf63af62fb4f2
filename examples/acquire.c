/* examples/acquire.c -- the C ABI without any wrapper: acquire one recording window.
 *
 *   gcc -Iinclude examples/acquire.c -Lassignment-for-aae6102_gnss-sdr_b200/gnssacq -lgnssacq -o acquire
 *   LD_LIBRARY_PATH=assignment-for-aae6102_gnss-sdr_b200/gnssacq ./acquire Opensky.bin 5000
 *
 * Does what SDR_main.m:17-23 does for the acquisition stage: initParameters defaults, seek to
 * skip*Sample*precision*type bytes (acquisition.m:27), read 20 ms, search 32 PRNs, refine the carrier of
 * the acquired ones with an 11 ms block (acquisition.m:89-121), print the lines acquisition.m prints.
 */
#include <stdio.h>
#include <stdlib.h>
#include "gnssacq.h"

int main(int argc, char** argv) {
    gnssacq_config cfg;
    gnssacq_handle* h = NULL;
    gnssacq_result rows[GNSSACQ_MAX_PRN];
    int32_t sv[GNSSACQ_MAX_PRN], cp[GNSSACQ_MAX_PRN];
    double fine[GNSSACQ_MAX_PRN];
    const int L = 10;                                   /* acq.L, initParameters.m:55 */
    long skip_ms;
    size_t need, need_long, bytes_per_ms;
    unsigned char* buf;
    FILE* f;
    int rc, i, n = 0;

    if (argc < 2) { fprintf(stderr, "usage: %s <recording.bin> [skip_ms]\n", argv[0]); return 2; }
    skip_ms = argc > 2 ? atol(argv[2]) : 5000;          /* file.skip, initParameters.m:22 */
    gnssacq_config_default(&cfg);
    bytes_per_ms = (size_t)cfg.samples_per_ms * cfg.data_type * cfg.data_precision;
    need = gnssacq_if_bytes(&cfg);
    need_long = bytes_per_ms * (size_t)(L + 1);
    buf = (unsigned char*)malloc(need > need_long ? need : need_long);
    f = fopen(argv[1], "rb");
    if (!f || !buf) { perror("open"); return 1; }

    rc = gnssacq_create(&cfg, &h);
    if (rc) { fprintf(stderr, "gnssacq_create: %d %s\n", rc, gnssacq_last_error(NULL)); return 1; }

    fseek(f, (long)(skip_ms * (long)bytes_per_ms), SEEK_SET);
    if (fread(buf, 1, need, f) != need) { fprintf(stderr, "short read\n"); return 1; }
    printf("Acquiring... \n");
    rc = gnssacq_search(h, buf, need, rows, NULL);
    if (rc) { fprintf(stderr, "gnssacq_search: %d %s\n", rc, gnssacq_last_error(h)); return 1; }
    for (i = 0; i < cfg.n_prn; ++i)
        if (rows[i].acquired) {
            printf(" SV[%2d] SNR = %2.2f, Code phase = %5d, Raw Doppler = %5d \n", rows[i].prn, rows[i].snr_db,
                   rows[i].code_phase, (int)rows[i].doppler_hz);
            sv[n] = rows[i].prn;
            cp[n] = rows[i].code_phase;
            ++n;
        }
    if (!n) { printf("No satellites acquired. Check parameter settings ... \n"); return 0; }

    printf("Now refining Doppler freq... \n");
    fseek(f, (long)(skip_ms * (long)bytes_per_ms), SEEK_SET);
    if (fread(buf, 1, need_long, f) != need_long) { fprintf(stderr, "short read\n"); return 1; }
    rc = gnssacq_fine_frequency(h, buf, need_long, L, n, sv, cp, fine);
    if (rc) { fprintf(stderr, "gnssacq_fine_frequency: %d %s\n", rc, gnssacq_last_error(h)); return 1; }
    for (i = 0; i < n; ++i) printf(" SV[%2d] Fine Doppler = %5f \n", sv[i], fine[i] - cfg.if_hz);

    gnssacq_destroy(h);
    fclose(f);
    free(buf);
    return 0;
}
