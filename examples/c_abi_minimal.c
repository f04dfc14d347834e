/* Minimal C host for the library: the eggcarton run of examples/eggcarton/NS_eggcarton.py
 * (reference) through the C ABI only -- no Python, no torch.
 *
 *   gcc -std=c11 -Iinclude examples/c_abi_minimal.c -o c_abi_minimal \
 *       -Lgleipnir_b200 -lgleipnir_b200 -Wl,-rpath,$PWD/gleipnir_b200 -lm
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include "gleipnir_b200.h"

int main(void) {
    const int d = 2;
    const long long N = 10000;
    const double alpha = (double)N / (double)(N + 1);
    /* RemainingPriorMass(1e-6) in closed form (stopping_criterion.py:48-57) */
    const long long n_deaths = (long long)ceil(log(1e-6) / log(alpha));
    gl_prior priors[2] = {{GL_PRIOR_UNIFORM, 0, 0.0, 10.0 * 3.14159265358979323846},
                          {GL_PRIOR_UNIFORM, 0, 0.0, 10.0 * 3.14159265358979323846}};
    gl_model *model = NULL;
    if (gl_model_create_eggcarton(d, &model) != GL_OK) {
        fprintf(stderr, "model: %s\n", gl_last_error());
        return 1;
    }
    gl_ns_config cfg;
    cfg.population_size = N;
    cfg.batch_k = (int)(N / 10);
    cfg.iterations = 50;
    cfg.burn_in = 0;
    cfg.tuning_cycles = 2;
    cfg.proposal = GL_PROPOSAL_UNIFORM;
    cfg.walker_lanes = 0;
    cfg.seed = 1234;
    cfg.dead_capacity = n_deaths + N;
    cfg.row_offset = 0;
    cfg.global_population = N;
    double *dead = (double *)malloc(sizeof(double) * (size_t)(n_deaths + N) * (size_t)(d + 2));
    gl_ns_state st;
    if (gl_ns_run_host(&cfg, priors, d, model, n_deaths, dead, &st) != GL_OK) {
        fprintf(stderr, "run: %s\n", gl_last_error());
        return 1;
    }
    printf("ln Z = %.6f +- %.4f  (information %.4f, %lld dead rows, %lld logL evaluations)\n", st.log_z,
           sqrt(st.information / (double)N), st.information, (long long)st.n_rows, (long long)st.n_evals);
    printf("first dead point: log_l %.6f weight %.3e x = (%.4f, %.4f)\n", dead[0], dead[1], dead[2], dead[3]);
    free(dead);
    gl_model_destroy(model);
    return 0;
}
