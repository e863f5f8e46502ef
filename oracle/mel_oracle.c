/*
 * oracle/mel_oracle.c — TEST INFRASTRUCTURE ONLY (never linked into or called from the product path).
 *
 * CPU restatement, in plain C / f32, of the log-mel front end that norma reaches at
 *   /root/reference/src/models/whisper/model.rs:74   audio::pcm_to_mel(&config, data_slice, &mel_filters)
 * The arithmetic itself lives in the un-vendored third-party crate candle-transformers 0.7.2
 * (pinned at /root/reference/Cargo.lock:293-294, module `models::whisper::audio`); it is restated here from
 * its published algorithm: periodic Hann, un-centred 400/160 framing, recursive radix-2
 * FFT down to a naive 25-point DFT with twiddles evaluated in f32 at every butterfly, power spectrum with the
 * `p[j] += p[400-j]` fold, dense 4-way-unrolled mel projection, log10(max(.,1e-10)), and the global
 * `max(x, max-8)/4+1` normalisation over ALL frames including the 1500 zero-pad frames.
 *
 * PARITY UNPINNED: the reference's own tests hold no golden vector for this path; the only
 * pinned artefacts are the two filterbank byte files and candle's two shape known-answer tests, both of
 * which tests/test_oracle_mel.py checks.  Numerical cross-check is against an independent fp64 STFT.
 *
 * Threading mirrors candle (thread t handles frames t, t+T, ...; T = clamp(even(n_cpu), 2, 12)); the result
 * does not depend on T because every frame is written by exactly one thread.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <stddef.h>

#define CHUNK_LENGTH 30

static const float TWO_PI_F = 3.14159265358979323846f + 3.14159265358979323846f;

/* naive DFT, candle `audio::dft`: out interleaved re/im, length 2n */
static void dft_f32(const float *inp, size_t n, float *out) {
    float n_t = (float)n;
    for (size_t k = 0; k < n; ++k) {
        float k_t = (float)k;
        float re = 0.0f, im = 0.0f;
        for (size_t j = 0; j < n; ++j) {
            float j_t = (float)j;
            float angle = TWO_PI_F * k_t * j_t / n_t;
            re += inp[j] * cosf(angle);
            im -= inp[j] * sinf(angle);
        }
        out[2 * k] = re;
        out[2 * k + 1] = im;
    }
}

/* recursive radix-2 DIT, candle `audio::fft` */
static void fft_f32(const float *inp, size_t n, float *out) {
    if (n == 1) {
        out[0] = inp[0];
        out[1] = 0.0f;
        return;
    }
    if (n % 2 == 1) {
        dft_f32(inp, n, out);
        return;
    }
    size_t h = n / 2;
    float *even = (float *)malloc(sizeof(float) * h);
    float *odd = (float *)malloc(sizeof(float) * h);
    float *even_fft = (float *)malloc(sizeof(float) * 2 * h);
    float *odd_fft = (float *)malloc(sizeof(float) * 2 * h);
    for (size_t i = 0; i < n; ++i) {
        if (i % 2 == 0) even[i / 2] = inp[i];
        else odd[i / 2] = inp[i];
    }
    fft_f32(even, h, even_fft);
    fft_f32(odd, h, odd_fft);
    float n_t = (float)n;
    for (size_t k = 0; k < h; ++k) {
        float k_t = (float)k;
        float theta = TWO_PI_F * k_t / n_t;
        float re = cosf(theta);
        float im = -sinf(theta);
        float re_odd = odd_fft[2 * k];
        float im_odd = odd_fft[2 * k + 1];
        out[2 * k] = even_fft[2 * k] + re * re_odd - im * im_odd;
        out[2 * k + 1] = even_fft[2 * k + 1] + re * im_odd + im * re_odd;
        out[2 * (k + h)] = even_fft[2 * k] - re * re_odd + im * im_odd;
        out[2 * (k + h) + 1] = even_fft[2 * k + 1] - re * im_odd - im * re_odd;
    }
    free(even); free(odd); free(even_fft); free(odd_fft);
}

typedef struct {
    size_t ith, n_threads;
    const float *hann, *samples, *filters;
    size_t n_samples, fft_size, fft_step, n_len, n_mel;
    float *mel; /* shared output; frames are disjoint across threads */
} worker_args;

/* candle `log_mel_spectrogram_w` (speed_up = false) */
static void *mel_worker(void *p) {
    worker_args *a = (worker_args *)p;
    size_t fft_size = a->fft_size, n_fft = 1 + fft_size / 2;
    float *fft_in = (float *)calloc(fft_size, sizeof(float));
    float *fft_out = (float *)malloc(sizeof(float) * 2 * fft_size);
    size_t end = a->n_samples / a->fft_step + 1;
    if (end > a->n_len) end = a->n_len;
    for (size_t i = a->ith; i < end; i += a->n_threads) {
        size_t offset = i * a->fft_step;
        size_t avail = a->n_samples - offset;
        size_t lim = avail < fft_size ? avail : fft_size;
        for (size_t j = 0; j < lim; ++j) fft_in[j] = a->hann[j] * a->samples[offset + j];
        for (size_t j = lim; j < fft_size; ++j) fft_in[j] = 0.0f;
        fft_f32(fft_in, fft_size, fft_out);
        for (size_t j = 0; j < fft_size; ++j)
            fft_out[j] = fft_out[2 * j] * fft_out[2 * j] + fft_out[2 * j + 1] * fft_out[2 * j + 1];
        for (size_t j = 1; j < fft_size / 2; ++j) fft_out[j] += fft_out[fft_size - j];
        for (size_t j = 0; j < a->n_mel; ++j) {
            const float *f = a->filters + j * n_fft;
            float sum = 0.0f;
            size_t k = 0;
            size_t lim4 = n_fft >= 3 ? n_fft - 3 : 0;
            while (k < lim4) {
                sum += fft_out[k] * f[k] + fft_out[k + 1] * f[k + 1] + fft_out[k + 2] * f[k + 2] +
                       fft_out[k + 3] * f[k + 3];
                k += 4;
            }
            while (k < n_fft) {
                sum += fft_out[k] * f[k];
                k += 1;
            }
            float v = sum > 1e-10f ? sum : 1e-10f;
            a->mel[j * a->n_len + i] = log10f(v);
        }
    }
    free(fft_in); free(fft_out);
    return NULL;
}

/* candle `log_mel_spectrogram_`: returns n_len (frames) through *n_len_out; `out` must hold n_mel*n_len floats.
 * Call with out == NULL to query n_len only. */
int oracle_log_mel_spectrogram(const float *samples_in, size_t n_in, const float *filters, size_t fft_size,
                               size_t fft_step, size_t n_mel, int n_threads_req, float *out,
                               size_t *n_len_out) {
    size_t n_len = n_in / fft_step;
    size_t pad = 100 * CHUNK_LENGTH / 2;
    if (n_len % pad != 0) n_len = (n_len / pad + 1) * pad;
    n_len += pad;
    if (n_len_out) *n_len_out = n_len;
    if (!out) return 0;

    float *hann = (float *)malloc(sizeof(float) * fft_size);
    for (size_t i = 0; i < fft_size; ++i)
        hann[i] = 0.5f * (1.0f - cosf((TWO_PI_F * (float)i) / (float)fft_size));
    size_t n_samples = n_len * fft_step;
    float *samples = (float *)calloc(n_samples, sizeof(float));
    memcpy(samples, samples_in, sizeof(float) * n_in);

    int T = n_threads_req - n_threads_req % 2;
    if (T > 12) T = 12;
    if (T < 2) T = 2;

    memset(out, 0, sizeof(float) * n_len * n_mel);
    pthread_t th[12];
    worker_args args[12];
    for (int t = 0; t < T; ++t) {
        args[t] = (worker_args){(size_t)t, (size_t)T, hann, samples, filters, n_samples,
                                fft_size, fft_step, n_len, n_mel, out};
        pthread_create(&th[t], NULL, mel_worker, &args[t]);
    }
    for (int t = 0; t < T; ++t) pthread_join(th[t], NULL);

    size_t l = n_len * n_mel;
    float mmax = out[0];
    for (size_t i = 1; i < l; ++i)
        if (out[i] >= mmax) mmax = out[i];
    mmax -= 8.0f;
    for (size_t i = 0; i < l; ++i) {
        float v = out[i] > mmax ? out[i] : mmax;
        out[i] = v / 4.0f + 1.0f;
    }
    free(hann); free(samples);
    return 0;
}

/* candle `pcm_to_mel`: N_FFT = 400, HOP_LENGTH = 160 */
int oracle_pcm_to_mel(const float *pcm, size_t n, const float *filters, size_t n_mel, int n_threads,
                      float *out, size_t *n_len_out) {
    return oracle_log_mel_spectrogram(pcm, n, filters, 400, 160, n_mel, n_threads, out, n_len_out);
}

/* exposed for unit tests of the FFT restatement */
void oracle_fft(const float *inp, size_t n, float *out) { fft_f32(inp, n, out); }

/* candle `sinusoids(length, channels)` (models::whisper::model) in f32 with libm, as Rust's
 * f32::ln/exp/sin/cos resolve to on Linux: out [length][channels] = [sin(t*inv) | cos(t*inv)]. */
void oracle_sinusoids(size_t length, size_t channels, float *out) {
    size_t half = channels / 2;
    float max_timescale = 10000.0f;
    float inc = logf(max_timescale) / (float)(half - 1);
    for (size_t t = 0; t < length; ++t)
        for (size_t i = 0; i < half; ++i) {
            float inv = expf((float)i * (-inc));
            float a = (float)t * inv;
            out[t * channels + i] = sinf(a);
            out[t * channels + half + i] = cosf(a);
        }
}
