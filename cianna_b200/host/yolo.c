/*
 * yolo.c - YOLO output-layer set-up of the host library.
 *
 * Same call surface, defaults and messages as upstream set_yolo_params / set_yolo_activ
 * (src/activ_functions.c:970-1032, :1129-1477); the layer's copy of the parameters becomes a
 * cb200_yolo_desc plus three small device tables and one association workspace (include/cianna_b200.h)
 * instead of upstream's five per-cell scratch arrays.
 */
#include <math.h>
#include <string.h>
#include <time.h>
#include "cianna.h"

static const char *iou_names[4] = { "Classical IoU", "Generalized GIoU", "Distance DIoU", "Distance DIoU2" };
static const char *dist_names[3] = { "Prior dist. IoU", "Prior dist. SIZE", "Prior dist. OFFSET" };

/* per-IoU-flavour default thresholds: good-but-not-best, low-IoU re-association, min IoU to fit
 * prob / obj / class / param, then the two "difficult" limits (src/activ_functions.c:1313-1345) */
static const float default_IoU_limits[4][8] = {
	{ 0.5f,  0.1f,  0.0f,  0.0f,  0.2f,  0.2f, 0.5f, 0.3f },   /* IoU   */
	{ 0.4f, -0.5f, -1.0f, -1.0f, -0.3f, -0.3f, 0.4f, 0.2f },   /* GIoU  */
	{ 0.3f, -0.6f, -1.0f, -1.0f, -0.5f, -0.5f, 0.3f, 0.1f },   /* DIoU  */
	{ 0.3f, -0.5f, -1.0f, -1.0f, -0.4f, -0.4f, 0.3f, 0.1f },   /* DIoU2 */
};
static const float default_scales[6] = { 2.0f, 2.0f, 1.0f, 2.0f, 1.0f, 1.0f };
static const float default_slopes_and_maxes[6][3] = {
	{ 1.0f, 6.0f, -6.0f }, { 1.0f, 1.6f, -1.6f }, { 1.0f, 6.0f, -6.0f },
	{ 1.0f, 6.0f, -6.0f }, { 1.0f, 6.0f, -6.0f }, { 1.0f, 1.2f, -0.2f },
};

static void print_float_row(const char *head, const float *v, int n, int stride, const char *fmt)
{
	int i;
	printf("%s", head);
	for (i = 0; i < n; i++) printf(fmt, v[i * stride]);
	printf("]\n");
}

int set_yolo_params(network *net, size_t nb_box, int nb_class, int nb_param, int max_nb_obj_per_image, const char *IoU_type_char,
	const char *prior_dist_type_char, float *prior_size, float *yolo_noobj_prob_prior, int fit_dim,
	int strict_box_size, int rand_startup, float rand_prob_best_box_assoc, float rand_prob, float min_prior_forced_scaling, float *scale_tab,
	float **slopes_and_maxes_tab, float *param_ind_scale, float *IoU_limits, int *fit_parts, int class_softmax,
	int diff_flag, const char *error_type, int no_override, int raw_output)
{
	yolo_param *y = net->y_param;
	int i, j;

	y->no_override = no_override;
	y->raw_output = raw_output;
	if (y->fit_dim > 0) { printf("\n ERROR: Trying to update existing YOLO layer setup is not supported yet\n"); exit(EXIT_FAILURE); }
	if (max_nb_obj_per_image > 0 && (1 + max_nb_obj_per_image * (7 + nb_param + diff_flag)) != net->output_dim) {
		printf("\n ERROR: Network output dim (target) specified in init_network and YOLO's \"max_nb_obj_per_image\" values do not match.\n");
		printf(" Output_dim should be equal to 1+max_nb_obj_per_image*(7+nb_param).\n");
		printf(" Got output_dim = %d, and max_nb_obj_per_image = %d \n\n", net->output_dim, max_nb_obj_per_image);
		exit(EXIT_FAILURE);
	}
	if ((int)nb_box > CB200_YOLO_MAX_BOX || nb_box == 0) {
		printf("\n ERROR: the B200 core handles 1 to %d YOLO boxes per grid cell (got %d).\n", CB200_YOLO_MAX_BOX, (int)nb_box);
		exit(EXIT_FAILURE);
	}

	if (strcmp(IoU_type_char, "IoU") == 0) y->IoU_type = CB200_IOU;
	else if (strcmp(IoU_type_char, "GIoU") == 0) y->IoU_type = CB200_GIOU;
	else if (strcmp(IoU_type_char, "DIoU") == 0) y->IoU_type = CB200_DIOU;
	else if (strcmp(IoU_type_char, "DIoU2") == 0) y->IoU_type = CB200_DIOU2;
	else { printf("\n WARNING: Unrecognized IoU type: %s, fallback to default GIoU\n", IoU_type_char); y->IoU_type = CB200_GIOU; }

	if (strcmp(prior_dist_type_char, "IoU") == 0 || strcmp(prior_dist_type_char, "IOU") == 0) y->prior_dist_type = CB200_DIST_IOU;
	else if (strcmp(prior_dist_type_char, "SIZE") == 0) y->prior_dist_type = CB200_DIST_SIZE;
	else if (strcmp(prior_dist_type_char, "OFFSET") == 0) y->prior_dist_type = CB200_DIST_OFFSET;
	else { printf("\n WARNING: Unrecognized prior dist. type: %s, fallback to default dist. Size\n", prior_dist_type_char); y->prior_dist_type = CB200_DIST_SIZE; }

	y->strict_box_size_association = strict_box_size;
	y->rand_startup = rand_startup < 0 ? 64000 : rand_startup;
	y->rand_prob_best_box_assoc = rand_prob_best_box_assoc < 0.0f ? 0.0f : rand_prob_best_box_assoc;
	y->rand_prob = rand_prob < 0.0f ? 0.0f : rand_prob;
	y->min_prior_forced_scaling = min_prior_forced_scaling <= 0.0f ? 0.0f : min_prior_forced_scaling;
	y->fit_dim = fit_dim;

	y->nb_box = (int)nb_box; y->nb_class = nb_class; y->nb_param = nb_param;
	y->max_nb_obj_per_image = max_nb_obj_per_image;
	y->class_softmax = class_softmax; y->diff_flag = diff_flag;

	y->prior_size = (float *)calloc(3 * nb_box, sizeof(float));
	if (prior_size != NULL) memcpy(y->prior_size, prior_size, 3 * nb_box * sizeof(float));
	y->noobj_prob_prior = (float *)calloc(nb_box, sizeof(float));
	for (i = 0; i < (int)nb_box; i++) y->noobj_prob_prior[i] = yolo_noobj_prob_prior != NULL ? yolo_noobj_prob_prior[i] : 0.2f;
	y->param_ind_scale = (float *)calloc(nb_param > 0 ? nb_param : 1, sizeof(float));
	for (i = 0; i < nb_param; i++) y->param_ind_scale[i] = param_ind_scale != NULL ? param_ind_scale[i] : 1.0f;

	/* user tables only override the entries that are not their "unset" marker */
	for (i = 0; i < 6; i++) y->scale_tab[i] = (scale_tab != NULL && scale_tab[i] > 0.0f) ? scale_tab[i] : default_scales[i];
	for (i = 0; i < 6; i++) {
		for (j = 0; j < 3; j++) y->slopes_and_maxes_tab[i][j] = default_slopes_and_maxes[i][j];
		if (slopes_and_maxes_tab != NULL) {
			if (slopes_and_maxes_tab[i][0] > 0.0f) y->slopes_and_maxes_tab[i][0] = slopes_and_maxes_tab[i][0];
			if (slopes_and_maxes_tab[i][1] < 100000.0f) y->slopes_and_maxes_tab[i][1] = slopes_and_maxes_tab[i][1];
			if (slopes_and_maxes_tab[i][2] > -100000.0f) y->slopes_and_maxes_tab[i][2] = slopes_and_maxes_tab[i][2];
		}
	}
	for (i = 0; i < 8; i++) y->IoU_limits[i] = (IoU_limits != NULL && IoU_limits[i] > -1.99f) ? IoU_limits[i] : default_IoU_limits[y->IoU_type][i];
	for (i = 0; i < 6; i++) y->fit_parts[i] = 1;
	if (nb_class <= 0) y->fit_parts[4] = -1;
	if (nb_param <= 0) y->fit_parts[5] = -1;
	if (fit_parts != NULL)
		for (i = 0; i < 6; i++)
			if (fit_parts[i] > -2) y->fit_parts[i] = fit_parts[i];

	if (strcmp(error_type, "complete") == 0) y->error_type = CB200_ERR_COMPLETE;
	else if (strcmp(error_type, "natural") == 0) y->error_type = CB200_ERR_NATURAL;
	else { printf(" WARNING: Unrecognized YOLO display error type %s, fallback to default \"natural\"\n", error_type); y->error_type = CB200_ERR_NATURAL; }

	printf("\n YOLO layer setup \n -------------------------------------------------------------------\n");
	printf(" Nboxes = %d\n Nclasses = %d\n Nparams = %d\n IoU type = %s\n", y->nb_box, y->nb_class, y->nb_param, iou_names[y->IoU_type]);
	printf(" Classification type: %s\n", y->class_softmax ? "softmax-CrossEntropy" : "sigmoid-MSE");
	printf(" Nb dim fitted : %d\n\n", y->fit_dim);
	print_float_row(" W priors = [", y->prior_size + 0, y->nb_box, 3, "%4.4f ");
	print_float_row(" H priors = [", y->prior_size + 1, y->nb_box, 3, "%4.4f ");
	print_float_row(" D priors = [", y->prior_size + 2, y->nb_box, 3, "%4.4f ");
	print_float_row(" No obj. prob. priors\n          = [", y->noobj_prob_prior, y->nb_box, 1, "%4.4f ");
	printf(" Fit parts: (Pos., Size, Prob., Obj., Class., Param.)\n   = [");
	for (i = 0; i < 6; i++) printf(" %d ", y->fit_parts[i]);
	printf("]\n");
	print_float_row(" Error scales: (Pos., Size, Prob., Obj., Class., Param.)\n   = [", y->scale_tab, 6, 1, " %5.3f ");
	print_float_row(" IoU lim.: (GdNotBest, LowBest, Prob., Obj., Class., Param., diffIoUlim, diffObjlim)\n   = [", y->IoU_limits, 8, 1, "%7.3f ");
	if (y->nb_param > 0) print_float_row(" Individual param. error scaling: \n   = [", y->param_ind_scale, y->nb_param, 1, "%7.3f ");
	printf("\n Activation slopes and limits: \n   = ");
	for (i = 0; i < 6; i++) printf("[%6.2f %6.2f %6.2f]\n     ", y->slopes_and_maxes_tab[i][0], y->slopes_and_maxes_tab[i][1], y->slopes_and_maxes_tab[i][2]);
	printf("\n *** Other training hyper-parameters *** \n");
	if (y->strict_box_size_association > 0)
		printf("  Strict box size association is ENABLED\n  Strict association Nb. good priors = %d\n", y->strict_box_size_association);
	else
		printf("  Strict box size association is DISABLED\n");
	printf("  Startup random association Nb. item : %d\n", y->rand_startup);
	printf("  Proportion of forced best prior assoc.: %5.3f\n", y->rand_prob_best_box_assoc);
	printf("  Proportion of forced random prior assoc.: %5.3f\n", y->rand_prob);
	printf("  Forced smallest prior association scaling : %6.3f\n", y->min_prior_forced_scaling);
	printf("  Closest prior association type : %s\n", dist_names[y->prior_dist_type]);
	printf("  Difficult flag in use: %s\n", y->diff_flag ? "True" : "False");
	printf("  Display error type : %s\n", y->error_type == CB200_ERR_COMPLETE ? "COMPLETE" : "NATURAL");
	printf("\n -------------------------------------------------------------------\n\n");

	return y->nb_box * (8 + y->nb_class + y->nb_param);
}

/* attach a private copy of the network-level YOLO set-up to the (last) conv layer and build its device side */
void set_yolo_activ(layer *current)
{
	network *net = current->c_network;
	conv_param *c = (conv_param *)current->param;
	yolo_param *y = (yolo_param *)malloc(sizeof(yolo_param));
	cb200_yolo_desc *d;
	float *tables;
	int i, j, cells = c->nb_area[0] * c->nb_area[1];
	size_t n_tab;

	if (net->y_param->nb_box * (8 + net->y_param->nb_class + net->y_param->nb_param) != c->nb_filters) {
		printf("%d %d\n", net->y_param->nb_box * (8 + net->y_param->nb_class + net->y_param->nb_param), c->nb_filters);
		printf("ERROR: Nb filters size mismatch in YOLO dimensions!\n");
		exit(EXIT_FAILURE);
	}
	*y = *net->y_param;
	current->activ_param = y;
	/* priors are fixed pixel sizes, never below one pixel (src/activ_functions.c:999-1001) */
	y->prior_size = (float *)calloc(3 * y->nb_box, sizeof(float));
	for (i = 0; i < y->nb_box; i++)
		for (j = 0; j < 3; j++)
			y->prior_size[i * 3 + j] = fmaxf(1.0f, net->y_param->prior_size[i * 3 + j]);
	for (i = 0; i < 3; i++) y->cell_size[i] = net->in_dims[i] / c->nb_area[i];

	n_tab = (size_t)4 * y->nb_box + (y->nb_param > 0 ? y->nb_param : 1);
	tables = (float *)calloc(n_tab, sizeof(float));
	memcpy(tables, y->prior_size, 3 * y->nb_box * sizeof(float));
	memcpy(tables + 3 * y->nb_box, y->noobj_prob_prior, y->nb_box * sizeof(float));
	memcpy(tables + 4 * y->nb_box, y->param_ind_scale, y->nb_param * sizeof(float));
	CB_CHECK(cb200_malloc((void **)&y->dev_tables, n_tab * sizeof(float)));
	CB_CHECK(cb200_h2d(y->dev_tables, tables, n_tab * sizeof(float), NULL));
	CB_CHECK(cb200_stream_sync(NULL));
	free(tables);

	d = &y->desc;
	memset(d, 0, sizeof(*d));
	d->dtype = net->dtype; d->batch = net->batch_size; d->length = net->batch_size;
	d->grid_h = c->nb_area[1]; d->grid_w = c->nb_area[0];
	d->nb_box = y->nb_box; d->nb_class = y->nb_class; d->nb_param = y->nb_param;
	d->max_nb_obj = y->max_nb_obj_per_image;
	d->target_stride = net->output_dim;
	d->fit_dim = y->fit_dim;
	d->IoU_type = y->IoU_type; d->prior_dist_type = y->prior_dist_type; d->error_type = y->error_type;
	d->class_softmax = y->class_softmax; d->diff_flag = y->diff_flag;
	d->strict_box_size_association = y->strict_box_size_association;
	d->rand_startup = y->rand_startup;
	d->rand_prob_best_box_assoc = y->rand_prob_best_box_assoc;
	d->rand_prob = y->rand_prob;
	d->min_prior_forced_scaling = y->min_prior_forced_scaling;
	for (i = 0; i < 3; i++) d->cell_size[i] = y->cell_size[i];
	memcpy(d->scale_tab, y->scale_tab, sizeof(d->scale_tab));
	memcpy(d->slopes_and_maxes, y->slopes_and_maxes_tab, sizeof(d->slopes_and_maxes));
	memcpy(d->IoU_limits, y->IoU_limits, sizeof(d->IoU_limits));
	memcpy(d->fit_parts, y->fit_parts, sizeof(d->fit_parts));
	d->prior_size = y->dev_tables;
	d->noobj_prob_prior = y->dev_tables + 3 * y->nb_box;
	d->param_ind_scale = y->dev_tables + 4 * y->nb_box;

	CB_CHECK(cb200_malloc((void **)&y->workspace, cb200_yolo_workspace_bytes(d)));
	CB_CHECK(cb200_malloc((void **)&y->parts_dev, (size_t)net->batch_size * 6 * sizeof(float)));
	CB_CHECK(cb200_host_alloc((void **)&y->parts_host, (size_t)net->batch_size * 6 * sizeof(float)));
	CB_CHECK(cb200_malloc((void **)&y->monitor_dev, (size_t)net->batch_size * cells * y->nb_box * 2 * sizeof(float)));
	CB_CHECK(cb200_host_alloc((void **)&y->monitor_host, (size_t)net->batch_size * cells * y->nb_box * 2 * sizeof(float)));
	CB_CHECK(cb200_malloc((void **)&y->box_state_dev, (size_t)net->batch_size * cells * y->nb_box * sizeof(int)));
	y->seed = (unsigned long long)time(NULL);
	y->step = 0;
}
