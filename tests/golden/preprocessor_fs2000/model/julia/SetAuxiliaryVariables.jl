function set_auxiliary_variables!(y::Vector{<: Real}, x::Vector{<: Real}, params::Vector{<: Real})
@inbounds begin
    y[15]=y[3];
    y[16]=y[2];
end
end
