/*
 * ref_probe.c - TEST INFRASTRUCTURE ONLY (oracle side, never linked into the product).
 *
 * Compiled together with the UNMODIFIED reference sources (from /root/reference/src,
 * see build_ref.sh) into oracle/_ref/<variant>/CIANNA.so.  It gives the parity tests
 * a way to (1) drive the reference's own layer objects one mini-batch at a time with
 * explicit inputs/targets and (2) read every intermediate tensor the reference keeps
 * (outputs, deltas, weights, momentum buffers, pool argmax maps, group-norm stats).
 *
 * The step driver follows the body of the reference training loop
 * (src/auxil.c:1797-1849: forward over layers, output_deriv_error, backprop in
 * reverse) and of the loss monitor (src/auxil.c:1851-1917) without the dataset,
 * printing and timing code around it.
 */
#include "prototypes.h"

/* -------- bookkeeping -------------------------------------------------- */
int probe_nb_layers(int net_id) { return networks[net_id]->nb_layers; }
int probe_layer_type(int net_id, int l) { return networks[net_id]->net_layers[l]->type; }
int probe_layer_activ(int net_id, int l) { return networks[net_id]->net_layers[l]->activation_type; }
float probe_layer_bias(int net_id, int l) { return networks[net_id]->net_layers[l]->bias_value; }

/* out[0..15]: geometry, meaning depends on the layer type */
void probe_layer_geom(int net_id, int l, int *out)
{
	layer *cur = networks[net_id]->net_layers[l];
	int i;
	for(i = 0; i < 16; i++) out[i] = 0;
	switch(cur->type)
	{
		case CONV: {
			conv_param *p = (conv_param*) cur->param;
			out[0] = p->nb_filters; out[1] = p->flat_f_size; out[2] = p->TC_padding;
			out[3] = p->prev_depth;
			for(i = 0; i < 3; i++) { out[4+i] = p->prev_size[i]; out[7+i] = p->nb_area[i]; out[10+i] = p->f_size[i]; }
			out[13] = p->stride[0]; out[14] = p->padding[0]; out[15] = p->int_padding[0];
			break; }
		case POOL: {
			pool_param *p = (pool_param*) cur->param;
			out[0] = p->nb_maps; out[1] = p->pool_type; out[2] = p->global; out[3] = p->prev_depth;
			for(i = 0; i < 3; i++) { out[4+i] = p->prev_size[i]; out[7+i] = p->nb_area[i]; out[10+i] = p->p_size[i]; }
			out[13] = p->stride[0]; out[14] = p->padding[0];
			break; }
		case DENSE: {
			dense_param *p = (dense_param*) cur->param;
			out[0] = p->nb_neurons; out[1] = p->in_size;
			break; }
		case NORM: {
			norm_param *p = (norm_param*) cur->param;
			out[0] = p->n_dim; out[1] = p->group_size; out[2] = p->nb_group; out[3] = p->set_off;
			out[4] = p->dim_offset; out[5] = p->output_dim; out[6] = p->data_format;
			break; }
		case LRN: {
			lrn_param *p = (lrn_param*) cur->param;
			out[0] = p->n_dim; out[1] = p->range; out[4] = p->dim_offset; out[5] = p->output_dim;
			break; }
	}
}

/* what: 0 output, 1 delta_o, 2 weights (filters / dense weights), 3 update (momentum buffer),
 * 4 pool_map (int*), 5 gamma, 6 beta, 7 mean, 8 var, 9 d_gamma, 10 d_beta,
 * 11 gamma_update, 12 beta_update, 13 im2col_input, 14 lrn local_scale, 15 input,
 * 16 dropout_mask (conv / dense / pool: the 0/1 floats of the last forward pass that drew one) */
void* probe_ptr(int net_id, int l, int what)
{
	layer *cur = networks[net_id]->net_layers[l];
	if(what == 0) return cur->output;
	if(what == 1) return cur->delta_o;
	if(what == 15) return cur->input;
	switch(cur->type)
	{
		case CONV: {
			conv_param *p = (conv_param*) cur->param;
			if(what == 2) return p->filters;
			if(what == 3) return p->update;
			if(what == 13) return p->im2col_input;
			if(what == 16) return cur->dropout_rate > 0.01f ? p->dropout_mask : NULL;	/* (not allocated otherwise) */
			break; }
		case DENSE: {
			dense_param *p = (dense_param*) cur->param;
			if(what == 2) return p->weights;
			if(what == 3) return p->update;
			if(what == 16) return cur->dropout_rate > 0.01f ? p->dropout_mask : NULL;	/* (not allocated otherwise) */
			break; }
		case POOL: {
			pool_param *p = (pool_param*) cur->param;
			if(what == 4) return p->pool_map;
			if(what == 16) return cur->dropout_rate > 0.01f ? p->dropout_mask : NULL;	/* (not allocated otherwise) */
			break; }
		case NORM: {
			norm_param *p = (norm_param*) cur->param;
			if(what == 5) return p->gamma;
			if(what == 6) return p->beta;
			if(what == 7) return p->mean;
			if(what == 8) return p->var;
			if(what == 9) return p->d_gamma;
			if(what == 10) return p->d_beta;
			if(what == 11) return p->gamma_update;
			if(what == 12) return p->beta_update;
			break; }
		case LRN: {
			lrn_param *p = (lrn_param*) cur->param;
			if(what == 14) return p->local_scale;
			break; }
	}
	return NULL;
}

/* -------- single mini-batch driver ------------------------------------- */

/* input: [batch_size][input_dim+1] floats (bias slot last), reference dataset layout
 * (src/auxil.c:320-329); target: [batch_size][output_dim]. */
void probe_forward(int net_id, float *input, int length, int is_inference)
{
	network *net = networks[net_id];
	int k;
	net->input = input;
	net->length = length;
	net->is_inference = is_inference;
	net->inference_drop_mode = AVG_MODEL;
	for(k = 0; k < net->nb_layers; k++)
		net->net_layers[k]->forward(net->net_layers[k]);
}

void probe_backward(int net_id, float *target, float lr, float momentum, float weight_decay)
{
	network *net = networks[net_id];
	int k;
	net->target = target;
	net->learning_rate = lr;
	net->momentum = momentum;
	net->weight_decay = weight_decay;
	output_deriv_error(net->net_layers[net->nb_layers-1]);
	for(k = net->nb_layers-1; k >= 0; k--)
		net->net_layers[k]->backprop(net->net_layers[k]);
}

/* per-element loss of the last layer into err[batch_size*out_size] (caller allocates, zeroed here) */
void probe_loss(int net_id, float *target, float *err, int out_size)
{
	network *net = networks[net_id];
	int k;
	net->target = target;
	net->out_size = out_size;
	for(k = 0; k < net->batch_size*out_size; k++) err[k] = 0.0f;
	net->output_error = err;
	output_error(net->net_layers[net->nb_layers-1]);
}

void probe_set_frozen(int net_id, int l, int frozen) { networks[net_id]->net_layers[l]->frozen = frozen; }
void probe_reset(void)
{
	/* forget all networks so that a test can build a fresh one with id 0 (nothing is freed upstream either) */
	nb_networks = 0;
}

/* -------- YOLO output layer (src/activ_functions.c:970-2985) ------------ */
/* the association pass looks at iter * train.size to leave its random start-up phase */
void probe_set_iter(int net_id, int iter, int train_size)
{
	networks[net_id]->iter = iter;
	networks[net_id]->train.size = train_size;
}

/* what: 0 IoU_monitor [B*cells][nb_box][2], 1 box_locked (int) [B*cells][nb_box], 2 prior_size [nb_box][3],
 * 3 cell_size (int) [3] - arrays of the last layer's own yolo_param copy */
void* probe_yolo_ptr(int net_id, int what)
{
	network *net = networks[net_id];
	yolo_param *a = (yolo_param*) net->net_layers[net->nb_layers-1]->activ_param;
	if(a == NULL) return NULL;
	if(what == 0) return a->IoU_monitor;
	if(what == 1) return a->box_locked;
	if(what == 2) return a->prior_size;
	if(what == 3) return a->cell_size;
	return NULL;
}

/* kernel-level driving of the output layer: the caller writes raw values into the last layer's output buffer
 * (probe_ptr what=0), then runs only its activation / only its error signal */
void probe_last_activation(int net_id, int length)
{
	network *net = networks[net_id];
	layer *last = net->net_layers[net->nb_layers-1];
	net->length = length;
	last->activation(last);
}

void probe_last_deriv_error(int net_id, float *target, int length)
{
	network *net = networks[net_id];
	net->target = target;
	net->length = length;
	output_deriv_error(net->net_layers[net->nb_layers-1]);
}
