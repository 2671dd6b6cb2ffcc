/* arch_tex.c - LaTeX table of the network architecture.
 *
 * Same file content, column order and option flags as upstream's print_architecture_tex (src/auxil.c:872-1097,
 * Python keyword list src/python_module.c:997-998), so scripts calling cnn.print_arch_tex(...) keep producing the same
 * .tex; built here as a column table (one formatter per column, dispatched on the layer type) instead of one printf
 * block per layer type.  pdflatex is run afterwards when it is installed, as upstream does. */
#include "cianna.h"
#include <sys/stat.h>

enum { COL_IN_SIZE, COL_SIZE, COL_F_SIZE, COL_STRIDE, COL_PADDING, COL_IN_PADDING, COL_OUT_SIZE, COL_ACTIV, COL_BIAS,
       COL_DROPOUT, COL_PARAMS, NB_COL };

static const char *col_spec[NB_COL] = {
	"p{2.0cm}<{\\centering}", "p{1.6cm}<{\\centering}", "p{2.0cm}<{\\centering}", "p{1.2cm}<{\\centering}",
	"p{1.2cm}<{\\centering}", "p{1.2cm}<{\\centering}", "p{2.0cm}<{\\centering}", "p{1.2cm}",
	"p{0.8cm}<{\\centering}", "p{1.2cm}<{\\centering}", "p{1.4cm}<{\\centering}" };
static const char *col_title[NB_COL] = {
	"In. size", "N. filters", "F. size", "Stride", "Padding", "Intern. Pad.", "Out. size", "Activ.", "Bias",
	"Dropout", "N. param." };
static const char *type_label[5] = { "Conv", "Pool", "Dense", "Norm", "LRN" };   /* indexed by layer_type_enum */
/* the table shows the activation family only (src/activ_functions.c:233-254), indexed by activation_functions_enum */
static const char *activ_short[5] = { "RELU", "LOGI", "SMAX", "YOLO", "LIN" };

static void put_dims(FILE *f, const int *d, char sep) { fprintf(f, "& %d%c%d%c%d ", d[0], sep, d[1], sep, d[2]); }

/* spatial extent a normalisation layer sees: it has no geometry of its own, so it is read off the layer before it */
static void put_prev_geometry(FILE *f, layer *l, int want_output)
{
	layer *p = l->previous;
	if (p == NULL) return;
	if (p->type == CONV) {
		conv_param *c = (conv_param *)p->param;
		put_dims(f, want_output ? c->nb_area : c->prev_size, 'x');
	} else if (p->type == POOL) {
		pool_param *q = (pool_param *)p->param;
		put_dims(f, want_output ? q->nb_area : q->prev_size, 'x');
	}
}

static void put_cell(FILE *f, layer *l, int col)
{
	conv_param *c = (conv_param *)l->param;
	pool_param *p = (pool_param *)l->param;
	norm_param *n = (norm_param *)l->param;
	dense_param *d = (dense_param *)l->param;
	const int has_weights = (l->type == CONV || l->type == DENSE);
	const int is_norm = (l->type == NORM || l->type == LRN);

	switch (col) {
	case COL_IN_SIZE:
		if (l->type == CONV) put_dims(f, c->prev_size, 'x');
		else if (l->type == POOL) put_dims(f, p->prev_size, 'x');
		else if (l->type == DENSE) fprintf(f, "& %d", d->in_size);
		else put_prev_geometry(f, l, 0);
		return;
	case COL_SIZE:
		if (l->type == CONV) fprintf(f, "& %d ", c->nb_filters);
		else if (l->type == DENSE) fprintf(f, "& %d ", d->nb_neurons);
		else if (l->type == NORM) fprintf(f, "& N.Gr. %d ", n->nb_group);
		else if (l->type == LRN) fprintf(f, "& ch\\_range: %d", cb_lrn_range(l));
		else fprintf(f, "& ");
		return;
	case COL_F_SIZE:
		if (l->type == CONV) put_dims(f, c->f_size, 'x');
		else if (l->type == POOL) put_dims(f, p->p_size, 'x');
		else if (l->type == NORM) fprintf(f, "& Gr.Size %d ", n->group_size);
		else fprintf(f, "& ");
		return;
	case COL_STRIDE:
		if (l->type == CONV) put_dims(f, c->stride, ':');
		else if (l->type == POOL) put_dims(f, p->stride, ':');
		else fprintf(f, "& ");
		return;
	case COL_PADDING:
		if (l->type == CONV) put_dims(f, c->padding, ':');
		else if (l->type == POOL) put_dims(f, p->padding, ':');
		else if (l->type == NORM) fprintf(f, "& Off %d ", n->set_off);
		else fprintf(f, "& ");
		return;
	case COL_IN_PADDING:
		if (l->type == CONV) put_dims(f, c->int_padding, ':');
		else fprintf(f, "& ");
		return;
	case COL_OUT_SIZE:
		if (l->type == CONV) put_dims(f, c->nb_area, 'x');
		else if (l->type == POOL) put_dims(f, p->nb_area, 'x');
		else if (l->type == DENSE) fprintf(f, "& %d ", d->nb_neurons);
		else put_prev_geometry(f, l, 1);
		return;
	case COL_ACTIV:
		fprintf(f, "& %s ", activ_short[l->activation_type >= RELU && l->activation_type <= LINEAR ? l->activation_type : LINEAR]);
		return;
	case COL_BIAS:
		if (has_weights) fprintf(f, "& %0.2f ", l->bias_value);
		else fprintf(f, "& ");
		return;
	case COL_DROPOUT:
		if (is_norm) fprintf(f, "& ");
		else fprintf(f, "& %d\\%% ", (int)(l->dropout_rate * 100.0f));
		return;
	case COL_PARAMS:
		if (has_weights) fprintf(f, "& %d ", l->nb_params);
		else fprintf(f, "& ");
		return;
	}
}

void print_architecture_tex(network *net, const char *path, const char *file_name, int l_size, int l_in_size,
                            int l_f_size, int l_out_size, int l_stride, int l_padding, int l_in_padding,
                            int l_activation, int l_bias, int l_dropout, int l_param_count)
{
	int shown[NB_COL], per_type[5] = { 0, 0, 0, 0, 0 };
	char tex_name[1024], command[2400];
	struct stat st;
	FILE *f;
	int i, k;

	shown[COL_IN_SIZE] = l_in_size;       shown[COL_SIZE] = l_size;          shown[COL_F_SIZE] = l_f_size;
	shown[COL_STRIDE] = l_stride;         shown[COL_PADDING] = l_padding;    shown[COL_IN_PADDING] = l_in_padding;
	shown[COL_OUT_SIZE] = l_out_size;     shown[COL_ACTIV] = l_activation;   shown[COL_BIAS] = l_bias;
	shown[COL_DROPOUT] = l_dropout;       shown[COL_PARAMS] = l_param_count;

	if (stat(path, &st) == -1) mkdir(path, 0700);
	snprintf(tex_name, sizeof(tex_name), "%s%s.tex", path, file_name);
	f = fopen(tex_name, "w+");
	if (f == NULL) { printf("\nERROR: cannot open %s for writing\n", tex_name); exit(EXIT_FAILURE); }

	fprintf(f, "\\documentclass[border=2pt]{standalone}\n\\usepackage[utf8]{inputenc}\n\\usepackage{array}\n"
	           "\\renewcommand{\\arraystretch}{1.1}\n\\begin{document}\n\\centering\n\\begin{tabular}{");
	fprintf(f, "p{0.6cm}p{1.4cm}");
	for (k = 0; k < NB_COL; k++) if (shown[k]) fprintf(f, "%s", col_spec[k]);
	fprintf(f, "}\n\\hline\\noalign{\\smallskip}\n");
	fprintf(f, "Id. & Type ");
	for (k = 0; k < NB_COL; k++) if (shown[k]) fprintf(f, "& %s ", col_title[k]);
	fprintf(f, "\\\\\n\\hline\\noalign{\\smallskip}\n");

	for (i = 0; i < net->nb_layers; i++) {
		layer *l = net->net_layers[i];
		if (l->type < 0 || l->type > LRN) { printf("ERROR: Unrecognized layer type in architechture tex\n"); exit(EXIT_FAILURE); }
		per_type[l->type] += 1;
		fprintf(f, "%d & %s\\_%d ", i + 1, type_label[l->type], per_type[l->type]);
		for (k = 0; k < NB_COL; k++) if (shown[k]) put_cell(f, l, k);
		fprintf(f, "\\\\\n");
	}
	fprintf(f, "\n\\hline\\noalign{\\smallskip}\n\\end{tabular}\n\\end{document}\n");
	fclose(f);

	if (system("command -v pdflatex > /dev/null 2>&1") == 0) {
		snprintf(command, sizeof(command), "pdflatex --interaction=batchmode -output-directory=%s %s", path, tex_name);
		if (system(command) != 0) printf("WARNING: pdflatex failed on %s\n", tex_name);
	} else
		printf("pdflatex not found: %s written, no .pdf produced\n", tex_name);
}
