/*
 * activ.c - activation-string handling and weight initialisers of the host library.
 * Follows the behaviour of upstream src/activ_functions.c:232-374,580-610 (string <-> parameters,
 * per-activation default bias value) and src/initializers.c (Xavier / LeCun / plain normal+uniform on rand()).
 */
#include <math.h>
#include <string.h>
#include "cianna.h"

void load_activ_param(layer *current, const char *activ)
{
	if (activ == NULL) { current->activation_type = LINEAR; return; }
	if (strncmp(activ, "SMAX", 4) == 0) current->activation_type = SOFTMAX;
	else if (strncmp(activ, "LIN", 3) == 0) current->activation_type = LINEAR;
	else if (strncmp(activ, "LOGI", 4) == 0) current->activation_type = LOGISTIC;
	else if (strncmp(activ, "YOLO", 4) == 0) current->activation_type = YOLO;
	else if (strncmp(activ, "RELU", 4) == 0) current->activation_type = RELU;
	else current->activation_type = LINEAR;
}

/* fills current->activ and the activation's default bias value (RELU 0.1, LOGI -1, LIN/YOLO 0.5, SMAX 0.1) */
void set_activ_defaults(layer *current, const char *activ)
{
	const char *t;
	current->activ.type = current->activation_type;
	current->activ.leak = 0.0f;
	current->activ.saturation = 0.0f;
	current->activ.beta = 0.0f;
	switch (current->activation_type) {
	case RELU:
		current->activ.saturation = 800.0f;
		current->activ.leak = 0.05f;
		current->bias_value = 0.1f;
		if (activ != NULL) {
			if ((t = strstr(activ, "_S")) != NULL) sscanf(t, "_S%f", &current->activ.saturation);
			if ((t = strstr(activ, "_L")) != NULL) sscanf(t, "_L%f", &current->activ.leak);
		}
		break;
	case LOGISTIC:
		current->activ.saturation = 6.0f;
		current->activ.beta = 1.0f;
		current->bias_value = -1.0f;
		if (activ != NULL) {
			if ((t = strstr(activ, "_S")) != NULL) sscanf(t, "_S%f", &current->activ.saturation);
			if ((t = strstr(activ, "_B")) != NULL) sscanf(t, "_B%f", &current->activ.beta);
		}
		break;
	case SOFTMAX:
		current->bias_value = 0.1f;
		break;
	case YOLO:
	case LINEAR:
	default:
		current->bias_value = 0.5f;
		break;
	}
}

void print_string_activ_param(layer *current, char *activ)
{
	switch (current->activation_type) {
	case LOGISTIC: sprintf(activ, "LOGI_S%0.2f_B%0.2f", current->activ.saturation, current->activ.beta); break;
	case SOFTMAX: sprintf(activ, "SMAX"); break;
	case YOLO: sprintf(activ, "YOLO"); break;
	case RELU: sprintf(activ, "RELU_S%0.2f_L%0.2f", current->activ.saturation, current->activ.leak); break;
	case LINEAR:
	default: sprintf(activ, "LIN"); break;
	}
}

void print_activ_param(FILE *f, layer *current, int f_bin)
{
	char temp[40];
	memset(temp, 0, sizeof(temp));
	print_string_activ_param(current, temp);
	if (f_bin) fwrite(temp, sizeof(char), 40, f);
	else fprintf(f, "%s ", temp);
}

/* ---- initialisers: Box-Muller on rand(), limits as upstream (src/initializers.c) ---- */
static double rnd_uniform(void) { return (double)rand() / ((double)RAND_MAX + 1.0); }
static double rnd_normal(void)
{
	double u1 = rnd_uniform(), u2 = rnd_uniform();
	if (u1 < 1e-300) u1 = 1e-300;
	return sqrt(-2.0 * log(u1)) * cos(2.0 * 3.14159265358979323846 * u2);
}

/* tab: [dim_out][dim_in] row-major for conv (dim_in = flat_f_size). Values drawn for every entry,
 * scale from fan-in / fan-out like Xavier (normal: sqrt(2/(in+out)), uniform: sqrt(6/(in+out))) or
 * LeCun (sqrt(1/in), sqrt(3/in)); Xavier uniform keeps upstream's sqrt(12/(in+out)) half-width; "normal"/"uniform" use init_scaling as sigma / half-width. */
void init_weights(float *tab, int dim_in, int dim_out, const char *init_fct, float init_scaling)
{
	size_t i, n = (size_t)dim_in * dim_out;
	double lim;
	int normal = 1;
	if (init_scaling < 0) init_scaling = 1.0f;
	if (init_fct == NULL || strcmp(init_fct, "xavier") == 0) { lim = sqrt(2.0 / (dim_in + dim_out)); normal = 1; }
	else if (strcmp(init_fct, "xavier_U") == 0) { lim = sqrt(12.0 / (dim_in + dim_out)); normal = 0; }
	else if (strcmp(init_fct, "lecun") == 0) { lim = sqrt(1.0 / dim_in); normal = 1; }
	else if (strcmp(init_fct, "lecun_U") == 0) { lim = sqrt(3.0 / dim_in); normal = 0; }
	else if (strcmp(init_fct, "normal") == 0) { lim = 1.0; normal = 1; }
	else if (strcmp(init_fct, "uniform") == 0) { lim = 1.0; normal = 0; }
	else { lim = sqrt(2.0 / (dim_in + dim_out)); normal = 1; }
	for (i = 0; i < n; i++)
		tab[i] = (float)(init_scaling * lim * (normal ? rnd_normal() : (2.0 * rnd_uniform() - 1.0)));
}
