# predict.rfsi_forest -- predict.ranger's regression call shapes
# (simulation/simulation.R:214; temp_croatia/2_predictions.R:167,174-176;
#  prcp_catalonia/5_prcp_prediction.R:287-291): returns a list with $predictions.
# A ranger object can be passed directly: it is imported on first use.
predict.rfsi_forest <- function(object, data, type = c("response", "quantiles"),
                                quantiles = c(0.1, 0.5, 0.9), device = 0L, ...) {
  type <- match.arg(type)
  vars <- object$independent.variable.names
  miss <- setdiff(vars, names(data))
  if (length(miss)) stop("Error: One or more independent variables not found in data: ", paste(miss, collapse = ", "))
  X <- data.matrix(as.data.frame(data)[, vars, drop = FALSE])   # by name; extra columns ignored
  storage.mode(X) <- "double"
  if (type == "response") {
    p <- .Call(C_rfsi_predict, object$handle, X, NULL, as.integer(device))
  } else {
    if (!object$quantreg) stop("Error: Set quantreg=TRUE in ranger(...) for quantile prediction.")
    p <- .Call(C_rfsi_predict, object$handle, X, as.double(quantiles), as.integer(device))
    colnames(p) <- paste("quantile=", quantiles)
  }
  list(predictions = p, num.trees = object$num.trees, num.independent.variables = length(vars),
       num.samples = nrow(X), treetype = "Regression")
}
predict_b200 <- function(object, data, ...) {
  if (inherits(object, "ranger")) object <- rfsi_import_ranger(object)
  predict.rfsi_forest(object, data, ...)
}
